// Persistent, warp-specialised tcgen05 GEMM for sm_100a:  C[M,N] = A[M,K] * B[N,K]^T  (bf16 in, fp32 accumulate).
//
//   A : activations / queries, row-major [M,K]  -> K-major UMMA operand A (TMEM lane = row of A)
//   B : nn.Linear weight [out,in] / corpus shard [n,D], row-major [N,K] -> K-major UMMA operand B
//
// Roles (256 threads, 1 CTA per SM):
//   warp 0      TMA producer   (one elected lane; kStages-deep smem ring of {A 128x64, B BNx64} bf16 tiles, SW128)
//   warp 1      MMA issuer     (one elected lane; tcgen05.mma M=128, N=BN, K=16; 2 TMEM accumulator stages)
//   warp 2      TMEM allocator
//   warps 4..7  epilogue       (tcgen05.ld 32 lanes x 32 columns per warp-instruction; functor `Epi` consumes them)
//
// The accumulator is double-buffered in TMEM so tile i's epilogue overlaps tile i+1's MMAs; smem stages and TMEM
// stages are handed over with mbarriers only (no __syncthreads in the main loop).
//
// `Epi` contract:
//   struct Epi { struct Params{...}; struct State{...};
//     static __device__ void init(State&, const Params&, int row_in_tile_lane);
//     static __device__ void tile(State&, const Params&, int m0, int n0, int row, uint32_t tmem_row_addr,
//                                 int M, int N);   // called per output tile; reads BN columns via tmem_ld_32x32
//     static __device__ void finish(State&, const Params&, int lane_row); }
#pragma once
#include "common.cuh"

namespace sgpt {

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 64;
constexpr int kGemmThreads = 256;

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kABytes = kGemmBM * kGemmBK * 2;
  static constexpr int kBBytes = BN * kGemmBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;  // 256 or 512 (power of two)
  // smem: [<=1024 align slack][stages * (A|B)][barriers + tmem holder]
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + 256;
};

template <int BN, class Epi>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, int M,
                    int N, int K, typename Epi::Params ep) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_tiles = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (M + kGemmBM - 1) / kGemmBM;
  const int n_tiles = (N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (K + kGemmBK - 1) / kGemmBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 4);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_holder, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * kGemmBM;
        const int n0 = (tile % n_tiles) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem_tiles + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_2d(sa, &tma_a, &full_bar[stage], kb * kGemmBK, m0);
          tma_load_2d(sb, &tma_b, &full_bar[stage], kb * kGemmBK, n0);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kGemmBM, BN, false);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem_tiles + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kABytes;
          const uint64_t da = make_smem_desc_sw128(sa, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < kGemmBK / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128-B swizzle row: +2 in the (addr >> 4) field
            umma_bf16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs above have read it
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ew = warp - 4;               // == warp % 4 -> TMEM lane quarter this warp may access
    const int row_in_tile = ew * 32 + lane;
    typename Epi::State st;
    Epi::init(st, ep, row_in_tile);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / n_tiles) * kGemmBM;
      const int n0 = (tile % n_tiles) * BN;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BN;
      Epi::template tile<BN>(st, ep, m0, n0, row_in_tile, trow, M, N);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    Epi::finish(st, ep, row_in_tile);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Epilogues
// ---------------------------------------------------------------------------------------------------------------
struct EpiNoState {};

// out_bf16[m, n] = act(acc + bias[n])          act = identity | gelu_new
template <bool kGelu>
struct EpiBiasActBF16 {
  struct Params {
    __nv_bfloat16* out;
    const float* bias;  // may be null
    int ldc;
  };
  using State = EpiNoState;
  static __device__ __forceinline__ void init(State&, const Params&, int) {}
  static __device__ __forceinline__ void finish(State&, const Params&, int) {}
  template <int BN>
  static __device__ __forceinline__ void tile(State&, const Params& p, int m0, int n0, int row, uint32_t trow,
                                              int M, int N) {
    const int m = m0 + row;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      const int n = n0 + c;
      if (n >= N) break;  // warp-uniform
      uint32_t v[32];
      tmem_ld_32x32(trow + c, v);
      tmem_ld_wait();
      if (m < M) {
        __nv_bfloat16* dst = p.out + static_cast<size_t>(m) * p.ldc + n;
        if (n + 32 <= N) {
          uint32_t o[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float a = __uint_as_float(v[2 * j]), b = __uint_as_float(v[2 * j + 1]);
            if (p.bias) { a += __ldg(p.bias + n + 2 * j); b += __ldg(p.bias + n + 2 * j + 1); }
            if (kGelu) { a = gelu_new(a); b = gelu_new(b); }
            o[j] = pack_bf16(a, b);
          }
          uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
          for (int j = 0; j < 4; ++j) d4[j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (n + j < N) {
              float a = __uint_as_float(v[j]);
              if (p.bias) a += __ldg(p.bias + n + j);
              if (kGelu) a = gelu_new(a);
              dst[j] = __float2bfloat16_rn(a);
            }
          }
        }
      }
    }
  }
};

// resid_f32[m, n] = resid_in[m, n] + acc + bias[n]      (fp32 residual stream; in-place allowed)
struct EpiResidualF32 {
  struct Params {
    float* out;
    const float* resid;  // may alias out
    const float* bias;   // may be null
    int ldc;
  };
  using State = EpiNoState;
  static __device__ __forceinline__ void init(State&, const Params&, int) {}
  static __device__ __forceinline__ void finish(State&, const Params&, int) {}
  template <int BN>
  static __device__ __forceinline__ void tile(State&, const Params& p, int m0, int n0, int row, uint32_t trow,
                                              int M, int N) {
    const int m = m0 + row;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      const int n = n0 + c;
      if (n >= N) break;
      uint32_t v[32];
      tmem_ld_32x32(trow + c, v);
      tmem_ld_wait();
      if (m < M) {
        const size_t off = static_cast<size_t>(m) * p.ldc + n;
        if (n + 32 <= N) {
          const float4* r4 = reinterpret_cast<const float4*>(p.resid + off);
          float4* o4 = reinterpret_cast<float4*>(p.out + off);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 r = r4[j];
            float b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
            if (p.bias) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n) + j);
              b0 = b.x; b1 = b.y; b2 = b.z; b3 = b.w;
            }
            r.x += __uint_as_float(v[4 * j + 0]) + b0;
            r.y += __uint_as_float(v[4 * j + 1]) + b1;
            r.z += __uint_as_float(v[4 * j + 2]) + b2;
            r.w += __uint_as_float(v[4 * j + 3]) + b3;
            o4[j] = r;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (n + j < N) {
              float a = __uint_as_float(v[j]) + p.resid[off + j];
              if (p.bias) a += __ldg(p.bias + n + j);
              p.out[off + j] = a;
            }
          }
        }
      }
    }
  }
};

// scores_f32[q, doc] = fixnan(acc * row_scale[q] * col_scale[doc])       (cos_sim / dot_score, NaN -> -1)
struct EpiScoresF32 {
  struct Params {
    float* out;
    const float* row_scale;  // per query  (1/||q|| for cos_sim; null = 1)
    const float* col_scale;  // per doc    (1/||d|| for cos_sim; null = 1)
    long long ldc;
  };
  using State = EpiNoState;
  static __device__ __forceinline__ void init(State&, const Params&, int) {}
  static __device__ __forceinline__ void finish(State&, const Params&, int) {}
  template <int BN>
  static __device__ __forceinline__ void tile(State&, const Params& p, int m0, int n0, int row, uint32_t trow,
                                              int M, int N) {
    const int m = m0 + row;
    const float rs = (p.row_scale && m < M) ? __ldg(p.row_scale + m) : 1.0f;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      const int n = n0 + c;
      if (n >= N) break;
      uint32_t v[32];
      tmem_ld_32x32(trow + c, v);
      tmem_ld_wait();
      if (m < M) {
        float* dst = p.out + static_cast<size_t>(m) * p.ldc + n;
        if (n + 32 <= N && (p.ldc & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 cs = make_float4(1.f, 1.f, 1.f, 1.f);
            if (p.col_scale) cs = __ldg(reinterpret_cast<const float4*>(p.col_scale + n) + j);
            float4 s;
            s.x = __uint_as_float(v[4 * j + 0]) * rs * cs.x;
            s.y = __uint_as_float(v[4 * j + 1]) * rs * cs.y;
            s.z = __uint_as_float(v[4 * j + 2]) * rs * cs.z;
            s.w = __uint_as_float(v[4 * j + 3]) * rs * cs.w;
            s.x = (s.x != s.x) ? -1.f : s.x;
            s.y = (s.y != s.y) ? -1.f : s.y;
            s.z = (s.z != s.z) ? -1.f : s.z;
            s.w = (s.w != s.w) ? -1.f : s.w;
            reinterpret_cast<float4*>(dst)[j] = s;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (n + j < N) {
              float s = __uint_as_float(v[j]) * rs * (p.col_scale ? __ldg(p.col_scale + n + j) : 1.f);
              dst[j] = (s != s) ? -1.f : s;
            }
          }
        }
      }
    }
  }
};

}  // namespace sgpt

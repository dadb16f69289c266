// S2/S3: exact row-wise top-k selection (torch.topk at XS:102-108; heapq.nlargest merge at XS:121-132).
//
// One CTA (1024 threads) per query.  MSD radix select on an order-preserving 32-bit image of the fp32 score
// (11 + 11 + 10 bits, warp-aggregated shared-memory histograms), then a gather of the k winners and an in-smem
// bitonic sort (score descending, id ascending).  The same kernel serves three sources through `TopkSrc`:
//   - a dense score row           scores[q, 0..n)                       (ids = id_base + column)
//   - a filtered candidate list   packed (score, local idx) pairs with a per-query count
//   - G gathered lists            [G, nq, L] scores + ids (cross-chunk / cross-shard merge; id < 0 = empty slot)
#include <math.h>

#include "../../include/sgpt_b200.h"
#include "common.cuh"
#include "host_utils.h"
#include "topk.cuh"

namespace sgpt {

__device__ __forceinline__ uint32_t score_key(float x) {
  const uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_score(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct Elem {
  uint32_t key;  // 0 = invalid / empty
  long long id;
};

__device__ __forceinline__ Elem load_elem(const TopkSrc& s, int q, int g, long long i) {
  Elem e;
  const long long off = g * s.stride_g + q * s.stride_q + i;
  if (s.packed != nullptr) {
    const uint2 p = s.packed[off];
    e.key = score_key(__uint_as_float(p.x));
    e.id = s.id_base + static_cast<long long>(p.y);
  } else {
    e.key = score_key(s.scores[off]);
    if (s.ids != nullptr) {
      e.id = s.ids[off];
      if (e.id < 0 || (s.exclude != nullptr && e.id == s.exclude[q])) e.key = 0;
    } else {
      e.id = s.id_base + i;
    }
  }
  return e;
}

__device__ __forceinline__ long long list_len(const TopkSrc& s, int q, int g) {
  if (s.counts != nullptr) {
    long long c = s.counts[g * s.nq + q];
    return c < s.L ? c : s.L;
  }
  return s.L;
}

constexpr int kTopkThreads = 1024;
constexpr int kBins = 2048;

// sorted[] / sorted_id[]: KP (power of two >= k) slots in dynamic smem
__global__ void __launch_bounds__(kTopkThreads) topk_select_kernel(TopkSrc src, int k, int KP,
                                                                   float* __restrict__ out_scores,
                                                                   long long* __restrict__ out_ids, TopkExtra extra) {
  extern __shared__ uint8_t dsm[];
  uint32_t* skey = reinterpret_cast<uint32_t*>(dsm);
  long long* sid = reinterpret_cast<long long*>(dsm + static_cast<size_t>(KP) * 4);
  __shared__ uint32_t hist[kBins];
  __shared__ uint32_t warp_tot[32];
  __shared__ uint32_t s_bin, s_need, s_cnt_gt, s_cnt_eq;
  __shared__ unsigned long long s_total;

  const int q = blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;

  // number of valid elements (needed to cap k)
  if (tid == 0) s_total = 0;
  __syncthreads();
  {
    unsigned long long local = 0;
    for (int g = 0; g < src.G; ++g) {
      const long long len = list_len(src, q, g);
      if (src.ids == nullptr) {
        if (tid == 0) local += static_cast<unsigned long long>(len);
      } else {
        for (long long i = tid; i < len; i += kTopkThreads) local += (load_elem(src, q, g, i).key != 0);
      }
    }
    if (local) atomicAdd(&s_total, local);
  }
  __syncthreads();
  const unsigned long long total = s_total;
  const uint32_t kk = static_cast<uint32_t>(total < static_cast<unsigned long long>(k) ? total : k);

  uint32_t prefix = 0, mask = 0, need = kk;
  if (kk > 0) {
    const int shifts[3] = {21, 10, 0};
    const uint32_t widths[3] = {11, 11, 10};
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
      const int shift = shifts[pass];
      const uint32_t bmask = (1u << widths[pass]) - 1u;
      for (int i = tid; i < kBins; i += kTopkThreads) hist[i] = 0;
      __syncthreads();
      for (int g = 0; g < src.G; ++g) {
        const long long len = list_len(src, q, g);
        const long long len_pad = (len + 31) & ~31ll;  // keep warps converged for match_any
        for (long long i = tid; i < len_pad; i += kTopkThreads) {
          uint32_t key = 0;
          bool ok = false;
          if (i < len) {
            key = load_elem(src, q, g, i).key;
            ok = (key != 0) && ((key & mask) == prefix);
          }
          const uint32_t bin = ok ? ((key >> shift) & bmask) : 0xffffffffu;
          const uint32_t peers = __match_any_sync(0xffffffffu, bin);
          if (ok && lane == (__ffs(peers) - 1)) atomicAdd(&hist[bin], __popc(peers));
        }
      }
      __syncthreads();
      // suffix scan over bins: thread t owns bins 2t, 2t+1
      const uint32_t h0 = hist[2 * tid], h1 = hist[2 * tid + 1];
      const uint32_t c = h0 + h1;
      uint32_t incl = c;  // inclusive suffix within the warp (higher lanes = higher bins)
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_down_sync(0xffffffffu, incl, o);
        if (lane + o < 32) incl += v;
      }
      if (lane == 0) warp_tot[warp] = incl;
      __syncthreads();
      uint32_t above = incl - c;  // elements in higher bins of this warp
      for (int w = warp + 1; w < 32; ++w) above += warp_tot[w];
      if (above < need && need <= above + c) {
        if (above + h1 >= need) {
          s_bin = 2 * tid + 1;
          s_need = need - above;
        } else {
          s_bin = 2 * tid;
          s_need = need - above - h1;
        }
      }
      __syncthreads();
      prefix |= s_bin << shift;
      mask |= bmask << shift;
      need = s_need;
      __syncthreads();
    }
  }
  // prefix == key of the kk-th largest element; `need` of the elements equal to it are winners.
  if (tid == 0) { s_cnt_gt = 0; s_cnt_eq = 0; }
  for (int i = tid; i < KP; i += kTopkThreads) { skey[i] = 0; sid[i] = -1; }
  __syncthreads();
  if (kk > 0) {
    const uint32_t n_gt = kk - need;
    for (int g = 0; g < src.G; ++g) {
      const long long len = list_len(src, q, g);
      for (long long i = tid; i < len; i += kTopkThreads) {
        const Elem e = load_elem(src, q, g, i);
        if (e.key == 0) continue;
        if (e.key > prefix) {
          const uint32_t slot = atomicAdd(&s_cnt_gt, 1u);
          skey[slot] = e.key;
          sid[slot] = e.id;
        } else if (e.key == prefix) {
          const uint32_t s = atomicAdd(&s_cnt_eq, 1u);
          if (s < need) {
            skey[n_gt + s] = e.key;
            sid[n_gt + s] = e.id;
          }
        }
      }
    }
  }
  __syncthreads();
  // bitonic sort, descending by key then ascending by id
  for (int size = 2; size <= KP; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < KP / 2; i += kTopkThreads) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const uint32_t ka = skey[lo], kb = skey[hi];
        const long long ia = sid[lo], ib = sid[hi];
        // "a before b" in the desired descending order
        const bool a_first = (ka > kb) || (ka == kb && ia <= ib);
        if (a_first != desc) {
          skey[lo] = kb; skey[hi] = ka;
          sid[lo] = ib; sid[hi] = ia;
        }
      }
      __syncthreads();
    }
  }
  if (out_scores != nullptr) {
    for (int i = tid; i < k; i += kTopkThreads) {
      const uint32_t key = skey[i];
      out_scores[static_cast<size_t>(q) * k + i] = key ? key_score(key) : -INFINITY;
      out_ids[static_cast<size_t>(q) * k + i] = key ? sid[i] : -1;
    }
  }
  // optional side outputs for the two-pass search: winners re-packed as the head of another candidate list, and the
  // k-th best score as that query's admission threshold
  if (extra.packed != nullptr) {
    for (int i = tid; i < static_cast<int>(kk); i += kTopkThreads)
      extra.packed[static_cast<long long>(q) * extra.cap + i] =
          make_uint2(__float_as_uint(key_score(skey[i])), static_cast<uint32_t>(sid[i] - src.id_base));
    if (tid == 0) extra.count[q] = static_cast<int>(kk);
  }
  if (extra.tau != nullptr && tid == 0)
    extra.tau[q] = (kk >= static_cast<uint32_t>(k)) ? key_score(skey[k - 1]) : -INFINITY;
}

int launch_topk_select(const TopkSrc& src, int nq, int k, float* out_scores, int64_t* out_ids, cudaStream_t stream,
                       const TopkExtra& extra) {
  if (k <= 0 || k > 4096) {
    set_error("top-k: k=%d outside [1, 4096]", k);
    return SGPT_ERR_INVALID;
  }
  int KP = 2;
  while (KP < k) KP <<= 1;
  const size_t dsm = static_cast<size_t>(KP) * 12;
  static bool attr_set = false;
  if (!attr_set) {
    SGPT_CHECK_CUDA(cudaFuncSetAttribute(topk_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 12));
    attr_set = true;
  }
  LaunchScope _ls(kCatTopk, stream);
  topk_select_kernel<<<nq, kTopkThreads, dsm, stream>>>(src, k, KP, out_scores, reinterpret_cast<long long*>(out_ids),
                                                        extra);
  SGPT_CHECK_CUDA(cudaGetLastError());
  return SGPT_OK;
}

}  // namespace sgpt

using namespace sgpt;

extern "C" int64_t sgpt_topk_workspace_bytes(int nq, int64_t n, int k) {
  (void)nq; (void)n; (void)k;
  return 256;  // selection runs entirely in shared memory; a token size keeps callers' allocation path uniform
}

extern "C" int sgpt_topk(const float* scores, int64_t lds, int nq, int64_t n, int k, int64_t id_base,
                         float* out_scores, int64_t* out_ids, void* ws, sgpt_stream_t stream_) {
  (void)ws;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGPT_REQUIRE(nq >= 0 && n >= 0 && lds >= n, "sgpt_topk: bad sizes nq=%d n=%lld lds=%lld", nq, (long long)n,
               (long long)lds);
  if (nq == 0) return SGPT_OK;
  TopkSrc src{};
  src.scores = scores;
  src.id_base = id_base;
  src.G = 1;
  src.nq = nq;
  src.L = n;
  src.stride_q = lds;
  return launch_topk_select(src, nq, k, out_scores, out_ids, stream);
}

extern "C" int sgpt_topk_merge(const float* in_scores, const int64_t* in_ids, int G, int nq, int k,
                               float* out_scores, int64_t* out_ids, const int64_t* exclude_ids, void* ws,
                               sgpt_stream_t stream_) {
  (void)ws;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGPT_REQUIRE(G >= 1 && nq >= 0, "sgpt_topk_merge: bad sizes G=%d nq=%d", G, nq);
  if (nq == 0) return SGPT_OK;
  TopkSrc src{};
  src.scores = in_scores;
  src.ids = reinterpret_cast<const long long*>(in_ids);
  src.exclude = reinterpret_cast<const long long*>(exclude_ids);
  src.G = G;
  src.nq = nq;
  src.L = k;
  src.stride_g = static_cast<long long>(nq) * k;
  src.stride_q = k;
  return launch_topk_select(src, nq, k, out_scores, out_ids, stream);
}

// S1+S2 for one corpus shard without materialising the [nq, n] score matrix.
//
// Two passes of the same tcgen05 similarity GEMM (corpus streamed from HBM exactly once in total):
//   pass A  scans a strided SAMPLE of the corpus tiles (every s-th 256-document tile, ~one tile per SM) and keeps all of
//           its scores; a radix select turns them into (i) the query's admission threshold tau[q] = its k-th best
//           sampled score and (ii) the k sampled winners, which seed the candidate lists (group 0).
//           tau[q] is a valid lower bound of the final k-th best score because k real documents already reach it.
//   pass B  scans all remaining tiles; the epilogue keeps only scores > tau[q] (expected k * n / n_sample survivors per
//           query) — the score matrix never reaches HBM.
//   final   radix select + sort over each query's candidate lists -> exact top-k.
// Candidates live in per-(query, group) lists, one group per (CTA, tile half) of the GEMM (gemm.cuh EpiFilterRows): no
// atomics, and every list is sized for the worst case (every score admitted), so no overflow path exists; only the
// touched prefix of each list ever generates memory traffic.  Small shards take the direct path (dense scores + select).
#include "../../include/sgpt_b200.h"
#include "gemm_api.h"
#include "host_utils.h"
#include "topk.cuh"

using namespace sgpt;

namespace {

inline int64_t padded_cols(int64_t n) { return (n + 3) & ~int64_t(3); }
inline int64_t align256(int64_t x) { return (x + 255) & ~int64_t(255); }

struct Plan {
  bool two_pass;
  int stride;          // tile stride of the sample
  FilterGeometry ga;   // pass-A lists: groups x L entries per query
  FilterGeometry gb;   // pass-B lists (group 0 of the block is the seed list, so gb.groups + 1 lists per query)
  int64_t stride_a, stride_b;  // entries per query block
};

Plan make_plan(int nq, int64_t n, int k) {
  Plan p{};
  const int64_t n_tiles = (n + kSimBN - 1) / kSimBN;
  const int sms = sm_count();
  // two-pass only pays off when the sample is a small fraction of the shard and can hold k documents
  p.two_pass = n_tiles >= 8ll * sms && static_cast<int64_t>(sms) * kSimBN >= 2ll * k && nq <= 128;
  if (!p.two_pass) return p;
  p.stride = static_cast<int>(n_tiles / sms);
  const int64_t sampled_tiles = (n_tiles + p.stride - 1) / p.stride;
  p.ga = filter_geometry(sampled_tiles);
  p.gb = filter_geometry(n_tiles - sampled_tiles);
  if (p.gb.L < k) p.gb.L = (k + 1) & ~1;  // the seed list must hold k entries
  p.stride_a = static_cast<int64_t>(p.ga.groups) * p.ga.L;
  p.stride_b = static_cast<int64_t>(p.gb.groups + 1) * p.gb.L;
  return p;
}

}  // namespace

constexpr int kQueryBlock = 128;  // queries per scan (one M-tile of the similarity GEMM)

extern "C" int64_t sgpt_search_workspace_bytes(int nq, int64_t n, int k) {
  if (nq > kQueryBlock) nq = kQueryBlock;  // larger batches are processed in blocks that reuse the workspace
  const Plan p = make_plan(nq, n, k);
  if (!p.two_pass) return static_cast<int64_t>(nq) * padded_cols(n) * 4 + 256;
  return align256(static_cast<int64_t>(nq) * p.stride_a * 8) + align256(static_cast<int64_t>(nq) * p.stride_b * 8) +
         align256(static_cast<int64_t>(p.ga.groups) * nq * 4) + align256(static_cast<int64_t>(p.gb.groups + 1) * nq * 4) +
         align256(nq * 4) + 256;
}

extern "C" int sgpt_search(const void* Q, const void* Cm, const float* q_scale, const float* c_scale, int nq,
                           int64_t n, int D, int k, int64_t id_base, float* out_scores, int64_t* out_ids, void* ws,
                           int64_t ws_bytes, sgpt_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGPT_REQUIRE(nq >= 0 && n >= 0 && k > 0, "sgpt_search: bad sizes nq=%d n=%lld k=%d", nq, (long long)n, k);
  SGPT_REQUIRE(n < (1ll << 31), "sgpt_search: shard too large for one launch (n=%lld)", (long long)n);
  SGPT_REQUIRE(D > 0 && D % 8 == 0, "sgpt_search: D=%d must be a positive multiple of 8", D);
  SGPT_REQUIRE(ws != nullptr && ws_bytes >= sgpt_search_workspace_bytes(nq, n, k),
               "sgpt_search: workspace too small (%lld < %lld bytes)", (long long)ws_bytes,
               (long long)sgpt_search_workspace_bytes(nq, n, k));
  if (nq == 0) return SGPT_OK;
  if (nq > kQueryBlock) {
    for (int q0 = 0; q0 < nq; q0 += kQueryBlock) {
      const int nb = (nq - q0 < kQueryBlock) ? nq - q0 : kQueryBlock;
      int rc = sgpt_search(static_cast<const uint8_t*>(Q) + static_cast<size_t>(q0) * D * 2, Cm,
                           q_scale ? q_scale + q0 : nullptr, c_scale, nb, n, D, k, id_base,
                           out_scores + static_cast<size_t>(q0) * k, out_ids + static_cast<size_t>(q0) * k, ws, ws_bytes,
                           stream_);
      if (rc != SGPT_OK) return rc;
    }
    return SGPT_OK;
  }
  const Plan p = make_plan(nq, n, k);
  if (!p.two_pass) {
    float* scores = static_cast<float*>(ws);
    const int64_t lds = padded_cols(n);
    int rc = sgpt_scores(Q, Cm, q_scale, c_scale, scores, lds, nq, n, D, stream);
    if (rc != SGPT_OK) return rc;
    return sgpt_topk(scores, lds, nq, n, k, id_base, out_scores, out_ids, nullptr, stream);
  }

  uint8_t* w = static_cast<uint8_t*>(ws);
  uint2* cand_a = reinterpret_cast<uint2*>(w);
  w += align256(static_cast<int64_t>(nq) * p.stride_a * 8);
  uint2* cand_b = reinterpret_cast<uint2*>(w);
  w += align256(static_cast<int64_t>(nq) * p.stride_b * 8);
  int* cnt_a = reinterpret_cast<int*>(w);
  const int64_t cnt_a_bytes = align256(static_cast<int64_t>(p.ga.groups) * nq * 4);
  w += cnt_a_bytes;
  int* cnt_b = reinterpret_cast<int*>(w);
  const int64_t cnt_b_bytes = align256(static_cast<int64_t>(p.gb.groups + 1) * nq * 4);
  w += cnt_b_bytes;
  float* tau = reinterpret_cast<float*>(w);

  // every group of a launch writes its own counts; the memset covers groups of CTAs that do not exist (tiny grids)
  SGPT_CHECK_CUDA(cudaMemsetAsync(cnt_a, 0, cnt_a_bytes + cnt_b_bytes, stream));
  // pass A: every score of the sampled tiles
  int rc = launch_filter_candidates(Q, Cm, q_scale, c_scale, nullptr, cand_a, cnt_a, p.stride_a, p.ga.L, 0, nq,
                                    static_cast<int>(n), D, /*tile_mode=*/1, p.stride, stream);
  if (rc != SGPT_OK) return rc;
  // thresholds + seed winners (group 0 of the pass-B lists)
  {
    TopkSrc src{};
    src.packed = cand_a;
    src.counts = cnt_a;
    src.G = p.ga.groups;
    src.nq = nq;
    src.L = p.ga.L;
    src.stride_g = p.ga.L;
    src.stride_q = p.stride_a;
    TopkExtra ex{};
    ex.packed = cand_b;
    ex.count = cnt_b;
    ex.cap = p.stride_b;
    ex.tau = tau;
    rc = launch_topk_select(src, nq, k, nullptr, nullptr, stream, ex);
    if (rc != SGPT_OK) return rc;
  }
  // pass B: the rest of the shard, admission threshold tau[q]
  rc = launch_filter_candidates(Q, Cm, q_scale, c_scale, tau, cand_b, cnt_b, p.stride_b, p.gb.L, 1, nq,
                                static_cast<int>(n), D, /*tile_mode=*/2, p.stride, stream);
  if (rc != SGPT_OK) return rc;
  // final exact selection over the candidates
  TopkSrc src{};
  src.packed = cand_b;
  src.counts = cnt_b;
  src.id_base = id_base;
  src.G = p.gb.groups + 1;
  src.nq = nq;
  src.L = p.gb.L;
  src.stride_g = p.gb.L;
  src.stride_q = p.stride_b;
  return launch_topk_select(src, nq, k, out_scores, out_ids, stream);
}

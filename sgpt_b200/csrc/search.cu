// S1+S2 for one corpus shard: scores (tcgen05 GEMM, corpus streamed once) -> exact top-k (radix select).
// v1 materialises the [nq, n] fp32 score panel in the caller's workspace between the two kernels.
#include "../../include/sgpt_b200.h"
#include "host_utils.h"

using namespace sgpt;

static inline int64_t padded_cols(int64_t n) { return (n + 3) & ~int64_t(3); }

extern "C" int64_t sgpt_search_workspace_bytes(int nq, int64_t n, int k) {
  (void)k;
  return static_cast<int64_t>(nq) * padded_cols(n) * 4 + 256;
}

extern "C" int sgpt_search(const void* Q, const void* Cm, const float* q_scale, const float* c_scale, int nq,
                           int64_t n, int D, int k, int64_t id_base, float* out_scores, int64_t* out_ids, void* ws,
                           int64_t ws_bytes, sgpt_stream_t stream) {
  SGPT_REQUIRE(nq >= 0 && n >= 0 && k > 0, "sgpt_search: bad sizes nq=%d n=%lld k=%d", nq, (long long)n, k);
  SGPT_REQUIRE(ws != nullptr && ws_bytes >= sgpt_search_workspace_bytes(nq, n, k),
               "sgpt_search: workspace too small (%lld < %lld bytes)", (long long)ws_bytes,
               (long long)sgpt_search_workspace_bytes(nq, n, k));
  if (nq == 0) return SGPT_OK;
  float* scores = static_cast<float*>(ws);
  const int64_t lds = padded_cols(n);
  int rc = sgpt_scores(Q, Cm, q_scale, c_scale, scores, lds, nq, n, D, stream);
  if (rc != SGPT_OK) return rc;
  return sgpt_topk(scores, lds, nq, n, k, id_base, out_scores, out_ids, nullptr, stream);
}

// Internal (C++) entry points of gemm.cu used by search.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sgpt {

constexpr int kSimBN = 256;  // corpus rows per similarity tile (N-tile width of the query x corpus GEMM)

// Geometry of the candidate lists the filter GEMM writes for a launch over `tiles` corpus tiles: every (CTA, column
// half) is a group with a private list of `L` entries per query (worst case: every score of every visited tile).
struct FilterGeometry {
  int groups;  // 2 x CTAs of the launch
  int L;       // entries per (query, group) list
};
FilterGeometry filter_geometry(long long tiles);

// Similarity GEMM whose epilogue appends every score > tau[q] (tau == nullptr: every score) as (score bits, local doc
// index) to cand[q * stride_q + (group0 + g) * L + ...] and writes counts[(group0 + g) * nq + q] for the groups g of this
// launch (filter_geometry of the visited tile count).  tile_mode/tile_stride select the corpus tiles visited
// (gemm.cuh TileMap).  nq <= 128.
int launch_filter_candidates(const void* Q, const void* C, const float* q_scale, const float* c_scale,
                             const float* tau, uint2* cand, int* counts, long long stride_q, int L, int group0, int nq,
                             int n, int D, int tile_mode, int tile_stride, cudaStream_t stream);

}  // namespace sgpt

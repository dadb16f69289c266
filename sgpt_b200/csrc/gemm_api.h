// Internal (C++) entry points of gemm.cu used by search.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sgpt {

constexpr int kSimBN = 256;  // corpus rows per similarity tile (N-tile width of the query x corpus GEMM)

// Similarity GEMM whose epilogue appends every score > tau[q] (tau == nullptr: every score) to cand[q][count[q]++]
// as (score bits, local doc index).  tile_mode/tile_stride select the corpus tiles visited (gemm.cuh TileMap).
int launch_filter_candidates(const void* Q, const void* C, const float* q_scale, const float* c_scale,
                             const float* tau, uint2* cand, int* count, long long cap, int nq, int n, int D,
                             int tile_mode, int tile_stride, cudaStream_t stream);

}  // namespace sgpt

// cholesky.cuh — interface of the tiled FP64 Cholesky (cholesky.cu) used by the BA engine.
#pragma once
#include <vector>

#include "cvb_internal.cuh"

namespace cvb_chol {

constexpr int T = 128;  // tile edge

// Tile-level structure of the factor: which 128x128 tiles of L are structurally non-zero, as launch lists.
struct TilePlan {
  int nt = 0;
  std::vector<int> h_col_ptr, h_row_idx;           // per tile column k: non-zero row tiles i > k
  std::vector<int> h_pair_ptr, h_pair_i, h_pair_j; // per tile column k: (i >= j) pairs of those rows (trailing updates)
  std::vector<int> h_pair_split;                   // per tile column k: how many of its pairs (listed first) lie in tile column k+1
  std::vector<int> h_rowc_ptr, h_rowc_idx;         // per tile row k: non-zero column tiles i < k (backward solve)
  std::vector<int> h_col_group;                    // optional, per tile column: id (>= 0) of an independent column group
                                                   // (its columns share no tile with other groups), -1 = main sequence
  int *d_row_idx = nullptr, *d_pair_i = nullptr, *d_pair_j = nullptr, *d_rowc_idx = nullptr;
  long n_tiles_L = 0;
  double flops = 0.0;   // flops of one numeric factorisation with this plan
  void build(int nt, std::vector<uint8_t> lower_mask);
  int upload(cvb_ctx* ctx, cudaStream_t st);
  void release();
};

// Extra streams/events of a factorisation (all events with timing disabled); nullptr → everything on one stream.
struct FactorStreams {
  cudaStream_t bulk = nullptr;       // low-priority stream of the bulk trailing updates (depth-1 lookahead)
  cudaEvent_t* ev = nullptr;         // 2 * nt events
  cudaStream_t group[8] = {};        // streams of the independent column groups
  int n_group = 0;
  cudaEvent_t fork = nullptr, join[8] = {};
};
int factor(cvb_ctx* ctx, double* S, int n_pad, double* linv, int* d_flag, const TilePlan& plan, cudaStream_t st,
           const FactorStreams* fs);
int solve(cvb_ctx* ctx, const double* L, int n_pad, const double* linv, double* b, double* tmp, double* x,
          const TilePlan& plan, cudaStream_t st, const FactorStreams* fs = nullptr);

}  // namespace cvb_chol

// Host interface of the in-tree FP64 dense linear algebra (dense_la.cu / dense_la.cuh): what the
// preconditioner set-up needs in place of a factorization library (ref: src/PoseGraph.cpp:598-613).
#pragma once
#include <cuda_runtime.h>

#include "dense_la.cuh"

namespace dpgo {
namespace dla {

struct SpdItem {
  double *A;   // n x n column-major, lower triangle valid on entry
  int n, lda;
};

// In place: lower triangle of every A <- lower triangle of A^-1 (symmetrize = true: the upper one too).
// Returns 0, or 1 + index of a matrix that is not positive definite, or -1 on a CUDA error
// (message through dla_last_error).  Blocks the host until the result is known.
int spd_inverse_batched(cudaStream_t st, const SpdItem *items, int count, bool symmetrize);

// C_b = alpha op(A_b) op(B_b) + beta C_b for every descriptor; queued on st, no host wait.
int gemm_batched(cudaStream_t st, const GemmDesc *descs, int count, const GemmFlags &flags);

const char *last_error();

}  // namespace dla
}  // namespace dpgo

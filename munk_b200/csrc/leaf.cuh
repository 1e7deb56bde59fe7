// One-CTA Cholesky + inverse of a diagonal block of order nb <= 128 (leaf.cu).
#pragma once
#include "common.cuh"

namespace munk {

// A (row-major, lda): lower triangle in, L out (zeros above the diagonal inside the block).
// Linv: 128 x 128 row-major, receives L^-1 (zeros outside the nb x nb lower triangle).
// *info (device): set to gidx + j + 1 if pivot j is the first non-positive one.
int launch_potrf_leaf(double* A, long lda, int nb, double* Linv, int gidx, int* info,
                      cudaStream_t stream);

}  // namespace munk

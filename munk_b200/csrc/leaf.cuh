// One-CTA Cholesky + inverse of a diagonal block of order nb <= 128 (leaf.cu).
#pragma once
#include "common.cuh"

namespace munk {

// A (row-major, lda): lower triangle in, L out (zeros above the diagonal inside the block).
// Linv: 128 x 128 row-major, receives L^-1 (zeros outside the nb x nb lower triangle).
// *info (device): set to gidx + j + 1 if pivot j is the first non-positive one.
int launch_potrf_leaf(double* A, long lda, int nb, double* Linv, int gidx, int* info,
                      cudaStream_t stream);

// Cholesky of a w x w diagonal block, w = 256 / 384 / 512, in one launch of co-operating CTAs
// (diag.cu); outputs as launch_potrf_leaf per 128-column step.  `counter`: a device word for the
// kernel's grid barrier (zeroed on `stream` by the call).
int launch_potrf_diag(double* A, long lda, int w, double* linv, int gidx, int* info, unsigned* counter,
                      cudaStream_t stream);

// B (m x s, in place) <- B L^-T in one launch (trsm_strip.cu): L is s x s lower triangular,
// s <= 512, `linv` holds the inverses of its 128 x 128 diagonal blocks (block k at linv + k*128*128).
int launch_trsm_strip(double* B, long ldb, int m, const double* L, long ldl, int s, const double* linv,
                      cudaStream_t stream);

}  // namespace munk

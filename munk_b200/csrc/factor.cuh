// Internal interface of the single-GPU factorisations (factor.cu).
#pragma once
#include "common.cuh"

namespace munk {

struct Workspace;

// bytes of leaf-inverse storage a matrix of order n needs (one 128x128 block per 128 columns)
size_t linv_bytes_for(int n);

// In-place lower Cholesky of the row-major SPD matrix A (only the lower triangle is read).
// On return the lower triangle holds L, the 128x128 diagonal blocks have zeros above the
// diagonal, `linv` holds the inverses of the diagonal blocks; ws->info is non-zero if a
// pivot was not positive.  Blocks above the diagonal keep whatever they held.
int potrf_lower(Workspace* ws, int n, double* A, long ld, double* linv, cudaStream_t stream);

// Two independent factorisations (two workspaces) with their block columns interleaved on one
// trailing-update stream; results and flags as two potrf_lower calls.
int potrf_pair_lower(Workspace* wa, int na, double* A, long lda, double* linva, Workspace* wb, int nb,
                     double* B, long ldb, double* linvb, cudaStream_t stream);

// In-place inverse of the lower-triangular factor produced by potrf_lower (same linv).
int trtri_lower(Workspace* ws, int n, double* L, long ld, double* linv, cudaStream_t stream);

// B (n x nrhs, row-major, in place) <- L^-1 B (transposed = 0) or L^-T B (transposed = 1), with
// the factor and leaf inverses left by potrf_lower.
// `stair` (device, optional, forward solve only): one int per 128 columns of B = first row where
// that column tile is non-zero (multiple of 128 not required); work above it is skipped.
int trsm_left(Workspace* ws, int n, const double* L, long ld, const double* linv, int transposed,
              int nrhs, double* B, long ldb, const int* stair, cudaStream_t stream);

// B (m x s, in place) <- B L^-T with the s x s factor L and its leaf inverses.
int trsm_right(Workspace* ws, int m, double* B, long ldb, const double* L, long ldl, int s,
               const double* linv, cudaStream_t stream);

}  // namespace munk

// Internal interface of the HBM-bound helper kernels (assemble.cu, scores.cu).
#pragma once
#include "common.cuh"

namespace munk {

// lval/ldiag == nullptr: unweighted simple graph (entries -1, integer degrees counted on the
// device); otherwise the off-diagonal / diagonal Laplacian entries themselves.  *bad (device)
// receives 1 + index of the first edge with an endpoint outside [0, n); such edges are skipped.
int assemble_reglap(int n, int nnz, const int* u, const int* v, const double* lval,
                    const double* ldiag, double lam, double* A, long ld, int* deg_scratch, int* bad,
                    cudaStream_t stream);
int gather_cols_lower(int n, const double* W, long ldw, int k, const int* idx, double* Q, long ldq,
                      cudaStream_t stream);
int transpose_lower(int n, const double* in, long ldi, double* out, long ldo, cudaStream_t stream);
int transpose_general(int rows, int cols, const double* in, long ldi, double* out, long ldo,
                      cudaStream_t stream);
int copy_lower(int n, const double* in, long ldi, double* out, long ldo, cudaStream_t stream);
int group_mean_rows(int ngroups, const int* gptr, const int* members, int ncols, const double* in,
                    long ldi, double* out, long ldo, cudaStream_t stream);
int gather_pairs(int npairs, const int* r, const int* c, const double* S, long lds, double* out,
                 cudaStream_t stream);

}  // namespace munk

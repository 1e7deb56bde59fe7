// Internal interface of the HBM-bound helper kernels (assemble.cu, scores.cu).
#pragma once
#include "common.cuh"

namespace munk {

int assemble_reglap(int n, int nnz, const int* u, const int* v, double lam, double* A, long ld,
                    int* deg_scratch, cudaStream_t stream);
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

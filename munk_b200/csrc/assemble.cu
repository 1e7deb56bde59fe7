// HBM-bound kernels around the factorisations: Laplacian assembly, gathers, transposes.
//
// assemble: A = I + lam * L for an unweighted simple graph, dense row-major FP64
//   (reference: munk/munk/munk.py:82-83, nx.laplacian_matrix(...).todense(); np.eye + lam*L).
//   Bit-exact with numpy: the diagonal is fl(1 + fl(lam*deg)) - multiply and add are kept
//   separate (__dmul_rn/__dadd_rn), an FMA would differ by 1 ulp for some degrees - and the
//   off-diagonal is -lam.
#include "common.cuh"
#include "kernels.cuh"

#include <algorithm>

namespace munk {

namespace {

// An edge whose endpoints are not both in [0, n) is skipped and reported (first one wins) in
// *bad as 1 + edge index: a caller of the C ABI cannot make the kernels write out of bounds.
__global__ void degree_kernel(int n, int nnz, const int* __restrict__ u, const int* __restrict__ v,
                              int* __restrict__ deg, int* __restrict__ bad) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < nnz) {
    const int a = u[e], b = v[e];
    if ((unsigned)a >= (unsigned)n || (unsigned)b >= (unsigned)n) {
      atomicCAS(bad, 0, e + 1);
      return;
    }
    atomicAdd(deg + a, 1);
    atomicAdd(deg + b, 1);
  }
}

// One pass over the whole n x ld array: zeros, and the diagonal entry of each row.
// Each thread stores 16 bytes; a block covers 4096 contiguous bytes of one row per step.
// The diagonal of L is the integer degree (`deg`) or, for weighted graphs, the float64 diagonal
// networkx computed (`ldiag`).
__global__ void __launch_bounds__(256)
fill_reglap_kernel(int n, long ld, double lam, const int* __restrict__ deg,
                   const double* __restrict__ ldiag, double* __restrict__ A) {
  const int row = blockIdx.x;
  const double diag = __dadd_rn(1.0, __dmul_rn(lam, ldiag ? ldiag[row] : (double)deg[row]));
  double2* __restrict__ p = reinterpret_cast<double2*>(A + (long)row * ld);
  const int npair = (int)(ld >> 1);
  const int dpair = row >> 1;
  for (int c = blockIdx.y * blockDim.x + threadIdx.x; c < npair; c += gridDim.y * blockDim.x) {
    double2 v = make_double2(0.0, 0.0);
    if (c == dpair) {
      if (row & 1) v.y = diag; else v.x = diag;
    }
    p[c] = v;
  }
}

// lval == nullptr: unweighted, L[a][b] = -1.  Otherwise lval[e] is the off-diagonal Laplacian
// entry itself (-weight, parallel edges already summed by the host).
__global__ void scatter_edges_kernel(int n, int nnz, const int* __restrict__ u, const int* __restrict__ v,
                                     const double* __restrict__ lval, double lam, long ld,
                                     double* __restrict__ A, int* __restrict__ bad) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < nnz) {
    const long a = u[e], b = v[e];
    if ((unsigned long)a >= (unsigned long)n || (unsigned long)b >= (unsigned long)n) {
      atomicCAS(bad, 0, e + 1);
      return;
    }
    const double w = __dmul_rn(lam, lval ? lval[e] : -1.0);
    A[a * ld + b] = w;
    A[b * ld + a] = w;
  }
}

// Q[i][l] = (i >= idx[l]) ? W[i][idx[l]] : 0   (columns of a lower-triangular matrix)
__global__ void gather_cols_lower_kernel(int n, const double* __restrict__ W, long ldw, int k,
                                         const int* __restrict__ idx, double* __restrict__ Q, long ldq) {
  const int l = blockIdx.y * blockDim.x + threadIdx.x;
  const int i = blockIdx.x;
  if (l < ldq) {
    double v = 0.0;
    if (l < k) {
      const int c = idx[l];
      if (i >= c) v = W[(long)i * ldw + c];
    }
    Q[(long)i * ldq + l] = v;
  }
}

// out[i][j] = (j >= i) ? in[j][i] : 0  — the upper-triangular transpose of a lower factor.
__global__ void transpose_lower_kernel(int n, const double* __restrict__ in, long ldi,
                                       double* __restrict__ out, long ldo) {
  __shared__ double tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;   // out tile origin: rows by.., cols bx..
  // read in[bx + ty][by + tx]
  for (int ty = threadIdx.y; ty < 32; ty += blockDim.y) {
    const int r = bx + ty, c = by + threadIdx.x;
    tile[ty][threadIdx.x] = (r < n && c < n && c <= r) ? in[(long)r * ldi + c] : 0.0;
  }
  __syncthreads();
  for (int ty = threadIdx.y; ty < 32; ty += blockDim.y) {
    const int r = by + ty, c = bx + threadIdx.x;
    if (r < n && c < n) out[(long)r * ldo + c] = tile[threadIdx.x][ty];
  }
}

// out[c][r] = in[r][c]
__global__ void transpose_kernel(int rows, int cols, const double* __restrict__ in, long ldi,
                                 double* __restrict__ out, long ldo) {
  __shared__ double tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int ty = threadIdx.y; ty < 32; ty += blockDim.y) {
    const int r = r0 + ty, c = c0 + threadIdx.x;
    tile[ty][threadIdx.x] = (r < rows && c < cols) ? in[(long)r * ldi + c] : 0.0;
  }
  __syncthreads();
  for (int ty = threadIdx.y; ty < 32; ty += blockDim.y) {
    const int c = c0 + ty, r = r0 + threadIdx.x;     // out row = c, out col = r
    if (c < cols && r < rows) out[(long)c * ldo + r] = tile[threadIdx.x][ty];
  }
}

__global__ void copy_lower_kernel(int n, const double* __restrict__ in, long ldi,
                                  double* __restrict__ out, long ldo) {
  const int r = blockIdx.x;
  for (int c = blockIdx.y * blockDim.x + threadIdx.x; c < n; c += gridDim.y * blockDim.x)
    out[(long)r * ldo + c] = (c <= r) ? in[(long)r * ldi + c] : 0.0;
}

// out[g][j] = mean over l in group g of in[members[l]][j]
__global__ void group_mean_rows_kernel(int ngroups, const int* __restrict__ gptr,
                                       const int* __restrict__ members, int ncols,
                                       const double* __restrict__ in, long ldi,
                                       double* __restrict__ out, long ldo) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int g = blockIdx.y;
  if (j >= ncols) return;
  const int lo = gptr[g], hi = gptr[g + 1];
  double s = in[(long)members[lo] * ldi + j];
  for (int l = lo + 1; l < hi; ++l) s += in[(long)members[l] * ldi + j];
  if (hi - lo > 1) s /= (double)(hi - lo);
  out[(long)g * ldo + j] = s;
}

__global__ void gather_pairs_kernel(int npairs, const int* __restrict__ r, const int* __restrict__ c,
                                    const double* __restrict__ S, long lds, double* __restrict__ out) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < npairs) out[l] = S[(long)r[l] * lds + c[l]];
}

}  // namespace

int assemble_reglap(int n, int nnz, const int* u, const int* v, const double* lval,
                    const double* ldiag, double lam, double* A, long ld, int* deg_scratch, int* bad,
                    cudaStream_t stream) {
  if (n <= 0) return MUNK_OK;
  if ((ld & 1) || ((uintptr_t)A & 15)) {
    set_last_error("assemble: A must be 16-byte aligned with an even leading dimension");
    return -7;
  }
  if (!ldiag) {
    MUNK_CUDA_TRY(cudaMemsetAsync(deg_scratch, 0, (size_t)n * sizeof(int), stream));
    if (nnz > 0) {
      degree_kernel<<<(nnz + 255) / 256, 256, 0, stream>>>(n, nnz, u, v, deg_scratch, bad);
      MUNK_CUDA_TRY(cudaGetLastError());
    }
  }
  const int npair = (int)(ld >> 1);
  dim3 grid(n, std::min(16, (npair + 255) / 256));
  fill_reglap_kernel<<<grid, 256, 0, stream>>>(n, ld, lam, deg_scratch, ldiag, A);
  MUNK_CUDA_TRY(cudaGetLastError());
  if (nnz > 0) {
    scatter_edges_kernel<<<(nnz + 255) / 256, 256, 0, stream>>>(n, nnz, u, v, lval, lam, ld, A, bad);
    MUNK_CUDA_TRY(cudaGetLastError());
  }
  return MUNK_OK;
}

int gather_cols_lower(int n, const double* W, long ldw, int k, const int* idx, double* Q, long ldq,
                      cudaStream_t stream) {
  if (n <= 0 || ldq <= 0) return MUNK_OK;
  dim3 grid(n, (unsigned)((ldq + 127) / 128));
  gather_cols_lower_kernel<<<grid, 128, 0, stream>>>(n, W, ldw, k, idx, Q, ldq);
  MUNK_CUDA_TRY(cudaGetLastError());
  return MUNK_OK;
}

int transpose_lower(int n, const double* in, long ldi, double* out, long ldo, cudaStream_t stream) {
  if (n <= 0) return MUNK_OK;
  dim3 grid((n + 31) / 32, (n + 31) / 32);
  transpose_lower_kernel<<<grid, dim3(32, 8), 0, stream>>>(n, in, ldi, out, ldo);
  MUNK_CUDA_TRY(cudaGetLastError());
  return MUNK_OK;
}

int transpose_general(int rows, int cols, const double* in, long ldi, double* out, long ldo,
                      cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return MUNK_OK;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32);
  transpose_kernel<<<grid, dim3(32, 8), 0, stream>>>(rows, cols, in, ldi, out, ldo);
  MUNK_CUDA_TRY(cudaGetLastError());
  return MUNK_OK;
}

int copy_lower(int n, const double* in, long ldi, double* out, long ldo, cudaStream_t stream) {
  if (n <= 0) return MUNK_OK;
  dim3 grid(n, std::min(16, (n + 255) / 256));
  copy_lower_kernel<<<grid, 256, 0, stream>>>(n, in, ldi, out, ldo);
  MUNK_CUDA_TRY(cudaGetLastError());
  return MUNK_OK;
}

int group_mean_rows(int ngroups, const int* gptr, const int* members, int ncols, const double* in,
                    long ldi, double* out, long ldo, cudaStream_t stream) {
  if (ngroups <= 0 || ncols <= 0) return MUNK_OK;
  dim3 grid((ncols + 255) / 256, ngroups);
  group_mean_rows_kernel<<<grid, 256, 0, stream>>>(ngroups, gptr, members, ncols, in, ldi, out, ldo);
  MUNK_CUDA_TRY(cudaGetLastError());
  return MUNK_OK;
}

int gather_pairs(int npairs, const int* r, const int* c, const double* S, long lds, double* out,
                 cudaStream_t stream) {
  if (npairs <= 0) return MUNK_OK;
  gather_pairs_kernel<<<(npairs + 255) / 256, 256, 0, stream>>>(npairs, r, c, S, lds, out);
  MUNK_CUDA_TRY(cudaGetLastError());
  return MUNK_OK;
}

}  // namespace munk

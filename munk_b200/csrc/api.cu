// extern "C" surface of libmunk_b200.so (declared in include/munk_b200.h).
#include "../../include/munk_b200.h"

#include "common.cuh"
#include "dgemm.cuh"
#include "factor.cuh"
#include "kernels.cuh"
#include "workspace.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstring>
#include <new>

struct munk_workspace : munk::Workspace {};

namespace munk {

static thread_local char g_err[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

}  // namespace munk

using namespace munk;

static inline cudaStream_t S(void* s) { return (cudaStream_t)s; }

extern "C" {

const char* munk_last_error_string(void) { return g_err; }
int munk_version(void) { return 100; }

int munk_workspace_create(int device, munk_workspace_t** out) {
  if (!out) return -2;
  *out = nullptr;
  MUNK_CUDA_TRY(cudaSetDevice(device));
  munk_workspace* ws = new (std::nothrow) munk_workspace();
  if (!ws) return -2;
  ws->device = device;
  cudaError_t e = cudaMalloc((void**)&ws->info, 8 * sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(ws->info, 0, 8 * sizeof(int));
  if (e == cudaSuccess) e = cudaMallocHost((void**)&ws->info_host, 8 * sizeof(int));
  if (e != cudaSuccess) {
    set_last_error("workspace allocation failed: %s", cudaGetErrorString(e));
    delete ws;
    return MUNK_ERR_CUDA + (int)e;
  }
  *out = ws;
  return MUNK_OK;
}

int munk_workspace_destroy(munk_workspace_t* ws) {
  if (!ws) return MUNK_OK;
  cudaSetDevice(ws->device);
  ws->release();
  delete ws;
  return MUNK_OK;
}

int munk_workspace_stats(munk_workspace_t* ws, long* launches, double* flops, int reset) {
  if (!ws) return -1;
  if (launches) *launches = ws->launches;
  if (flops) *flops = ws->flops;
  if (reset) { ws->launches = 0; ws->flops = 0.0; }
  return MUNK_OK;
}

size_t munk_linv_bytes(int n) { return linv_bytes_for(n); }

int munk_assemble_reglap(munk_workspace_t* ws, int n, int nnz, const int* u, const int* v,
                         double lam, double* A, long ld, void* stream) {
  if (!ws) return -1;
  if (n < 0) return -2;
  if (nnz < 0) return -3;
  if (nnz > 0 && (!u || !v)) return -4;
  if (ld < n) return -8;
  DeviceGuard guard(ws->device);
  // degree scratch: reuse the split-K buffer (never live at the same time on one stream)
  MUNK_TRY(ws->ensure_splitk(0, std::max(SPLITK_RESERVE, (size_t)(n + 2) * sizeof(int))));
  MUNK_TRY(assemble_reglap(n, nnz, u, v, nullptr, nullptr, lam, A, ld, (int*)ws->splitk[0], ws->info + 1,
                           S(stream)));
  ws->launches += (nnz > 0) ? 3 : 1;
  return MUNK_OK;
}

int munk_assemble_reglap_entries(munk_workspace_t* ws, int n, int nnz, const int* u, const int* v,
                                 const double* lval, const double* ldiag, double lam, double* A,
                                 long ld, void* stream) {
  if (!ws) return -1;
  if (n < 0) return -2;
  if (nnz < 0) return -3;
  if (nnz > 0 && (!u || !v || !lval)) return -4;
  if (n > 0 && !ldiag) return -7;
  if (ld < n) return -10;
  DeviceGuard guard(ws->device);
  MUNK_TRY(assemble_reglap(n, nnz, u, v, lval, ldiag, lam, A, ld, nullptr, ws->info + 1, S(stream)));
  ws->launches += (nnz > 0) ? 2 : 1;
  return MUNK_OK;
}

int munk_assemble_info(munk_workspace_t* ws, int* bad_edge_host, void* stream) {
  if (!ws || !bad_edge_host) return -1;
  DeviceGuard guard(ws->device);
  MUNK_CUDA_TRY(cudaMemcpyAsync(ws->info_host + 1, ws->info + 1, sizeof(int), cudaMemcpyDeviceToHost, S(stream)));
  MUNK_CUDA_TRY(cudaMemsetAsync(ws->info + 1, 0, sizeof(int), S(stream)));
  MUNK_CUDA_TRY(cudaStreamSynchronize(S(stream)));
  *bad_edge_host = ws->info_host[1];
  return MUNK_OK;
}

int munk_potrf(munk_workspace_t* ws, int n, double* A, long ld, double* linv, void* stream) {
  if (!ws) return -1;
  DeviceGuard guard(ws->device);
  if (n < 0) return -2;
  if (ld < n || (ld & 1)) return -4;
  if (!A || !linv) return -3;
  return potrf_lower(ws, n, A, ld, linv, S(stream));
}

int munk_potrf_pair(munk_workspace_t* wa, int na, double* A, long lda, double* linva, munk_workspace_t* wb,
                    int nb, double* B, long ldb, double* linvb, void* stream) {
  if (!wa || !wb) return -1;
  if (wa->device != wb->device) return -6;
  DeviceGuard guard(wa->device);
  if (na < 0) return -2;
  if (nb < 0) return -7;
  if (lda < na || (lda & 1)) return -4;
  if (ldb < nb || (ldb & 1)) return -9;
  if ((na > 0 && (!A || !linva)) || (nb > 0 && (!B || !linvb))) return -3;
  if (na == 0) return potrf_lower(wb, nb, B, ldb, linvb, S(stream));
  if (nb == 0) return potrf_lower(wa, na, A, lda, linva, S(stream));
  return potrf_pair_lower(wa, na, A, lda, linva, wb, nb, B, ldb, linvb, S(stream));
}

int munk_potrf_info(munk_workspace_t* ws, int* info_host, void* stream) {
  if (!ws || !info_host) return -1;
  DeviceGuard guard(ws->device);
  MUNK_CUDA_TRY(cudaMemcpyAsync(ws->info_host, ws->info, sizeof(int), cudaMemcpyDeviceToHost, S(stream)));
  MUNK_CUDA_TRY(cudaMemsetAsync(ws->info, 0, sizeof(int), S(stream)));
  MUNK_CUDA_TRY(cudaStreamSynchronize(S(stream)));
  *info_host = ws->info_host[0];
  return MUNK_OK;
}

int munk_potrf_info_async(munk_workspace_t* ws, int* info_dev, void* stream) {
  if (!ws || !info_dev) return -1;
  DeviceGuard guard(ws->device);
  MUNK_CUDA_TRY(cudaMemcpyAsync(info_dev, ws->info, sizeof(int), cudaMemcpyDeviceToDevice, S(stream)));
  MUNK_CUDA_TRY(cudaMemsetAsync(ws->info, 0, sizeof(int), S(stream)));
  return MUNK_OK;
}

int munk_trtri(munk_workspace_t* ws, int n, double* L, long ld, double* linv, void* stream) {
  if (!ws) return -1;
  DeviceGuard guard(ws->device);
  if (n < 0) return -2;
  if (ld < n || (ld & 1)) return -4;
  if (!L || !linv) return -3;
  return trtri_lower(ws, n, L, ld, linv, S(stream));
}

int munk_trsm_left(munk_workspace_t* ws, int n, const double* L, long ld, const double* linv,
                   int transposed, int nrhs, double* B, long ldb, const int* stair, void* stream) {
  if (!ws) return -1;
  DeviceGuard guard(ws->device);
  if (n < 0) return -2;
  if (ld < n || (ld & 1)) return -4;
  if (nrhs < 0) return -7;
  if (ldb < nrhs || (ldb & 1)) return -9;
  return trsm_left(ws, n, L, ld, linv, transposed, nrhs, B, ldb, stair, S(stream));
}

int munk_trsm_right(munk_workspace_t* ws, int m, double* B, long ldb, const double* L, long ldl,
                    int s, const double* linv, void* stream) {
  if (!ws) return -1;
  DeviceGuard guard(ws->device);
  if (m < 0 || s < 0) return -2;
  if (ldb < s || (ldb & 1) || ldl < s || (ldl & 1)) return -4;
  return trsm_right(ws, m, B, ldb, L, ldl, s, linv, S(stream));
}

int munk_dgemm_stair(munk_workspace_t* ws, int a_layout, int b_layout, int M, int N, int K,
                     double alpha, const double* A, long lda, const double* B, long ldb, double beta,
                     double* C, long ldc, int flags, const int* stair_m, const int* stair_n,
                     int row0, int k0, void* stream) {
  if (!ws) return -1;
  DeviceGuard guard(ws->device);
  if ((a_layout | b_layout) & ~1) return -2;
  if (M < 0 || N < 0 || K < 0) return -4;
  GemmDesc g{};
  g.M = M; g.N = N; g.K = K;
  g.A = A; g.lda = lda; g.a_layout = a_layout;
  g.B = B; g.ldb = ldb; g.b_layout = b_layout;
  g.C = C; g.ldc = ldc;
  g.alpha = alpha; g.beta = beta; g.flags = flags;
  g.stair_m = stair_m; g.stair_n = stair_n; g.stair_row0 = row0; g.stair_k0 = k0;
  return gemm_launch(ws, g, S(stream));
}

int munk_transpose(int rows, int cols, const double* in, long ldi, double* out, long ldo,
                   void* stream) {
  return transpose_general(rows, cols, in, ldi, out, ldo, S(stream));
}

int munk_gram_lower(munk_workspace_t* ws, int n, const double* W, long ldw, double* D, long ldd,
                    void* stream) {
  if (!ws) return -1;
  DeviceGuard guard(ws->device);
  GemmDesc g{};
  g.M = n; g.N = n; g.K = n;
  g.A = W; g.lda = ldw; g.a_layout = OP_MC;
  g.B = W; g.ldb = ldw; g.b_layout = OP_MC;
  g.C = D; g.ldc = ldd;
  g.alpha = 1.0; g.beta = 0.0;
  g.flags = GF_TILE_LOWER | GF_KLO_M | GF_KLO_N | GF_SYM_MIRROR;
  return gemm_launch(ws, g, S(stream));
}

int munk_dgemm(munk_workspace_t* ws, int a_layout, int b_layout, int M, int N, int K, double alpha,
               const double* A, long lda, const double* B, long ldb, double beta, double* C,
               long ldc, int flags, void* stream) {
  if (!ws) return -1;
  DeviceGuard guard(ws->device);
  if ((a_layout | b_layout) & ~1) return -2;
  if (M < 0 || N < 0 || K < 0) return -4;
  GemmDesc g{};
  g.M = M; g.N = N; g.K = K;
  g.A = A; g.lda = lda; g.a_layout = a_layout;
  g.B = B; g.ldb = ldb; g.b_layout = b_layout;
  g.C = C; g.ldc = ldc;
  g.alpha = alpha; g.beta = beta; g.flags = flags;
  return gemm_launch(ws, g, S(stream));
}

int munk_gather_cols_lower(int n, const double* W, long ldw, int k, const int* idx, double* Q,
                           long ldq, void* stream) {
  return gather_cols_lower(n, W, ldw, k, idx, Q, ldq, S(stream));
}

int munk_transpose_lower(int n, const double* W, long ldw, double* C, long ldc, void* stream) {
  return transpose_lower(n, W, ldw, C, ldc, S(stream));
}

int munk_download_upper(int n, const double* C, long ldc, double* host, long ldh, int row_block,
                        void* stream) {
  if (n < 0) return -1;
  if (!C && n > 0) return -2;
  if (ldc < n) return -3;
  if (!host && n > 0) return -4;
  if (ldh < n) return -5;
  if (row_block <= 0) return -6;
  for (long r0 = 0; r0 < n; r0 += row_block) {
    long rows = (n - r0 < row_block) ? (n - r0) : row_block;
    MUNK_CUDA_TRY(cudaMemcpy2DAsync(host + r0 * ldh + r0, (size_t)ldh * sizeof(double),
                                    C + r0 * ldc + r0, (size_t)ldc * sizeof(double),
                                    (size_t)(n - r0) * sizeof(double), (size_t)rows,
                                    cudaMemcpyDeviceToHost, S(stream)));
  }
  return MUNK_OK;
}

int munk_download_block(int rows, int cols, const double* C, long ldc, double* host, long ldh,
                        void* stream) {
  if (rows < 0) return -1;
  if (cols < 0) return -2;
  if (rows == 0 || cols == 0) return MUNK_OK;
  if (!C) return -3;
  if (ldc < cols) return -4;
  if (!host) return -5;
  if (ldh < cols) return -6;
  MUNK_CUDA_TRY(cudaMemcpy2DAsync(host, (size_t)ldh * sizeof(double), C, (size_t)ldc * sizeof(double),
                                  (size_t)cols * sizeof(double), (size_t)rows,
                                  cudaMemcpyDeviceToHost, S(stream)));
  return MUNK_OK;
}

int munk_copy_lower(int n, const double* in, long ldi, double* out, long ldo, void* stream) {
  return copy_lower(n, in, ldi, out, ldo, S(stream));
}

int munk_group_mean_rows(int ngroups, const int* gptr, const int* members, int ncols,
                         const double* in, long ldi, double* out, long ldo, void* stream) {
  return group_mean_rows(ngroups, gptr, members, ncols, in, ldi, out, ldo, S(stream));
}

int munk_gather_pairs(int npairs, const int* r, const int* c, const double* Sm, long lds,
                      double* out, void* stream) {
  return gather_pairs(npairs, r, c, Sm, lds, out, S(stream));
}

}  // extern "C"

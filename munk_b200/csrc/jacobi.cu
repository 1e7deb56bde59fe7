// Rank-revealing solve with a small symmetric positive semi-definite matrix:
//     Z = pinv(G) B,   G (k x k, k <= a few thousand),   B (k x n)
// the fallback behind np.linalg.pinv(source_C[source_idxs,:]) (munk/munk/munk.py:109) when the
// landmark Gram matrix G = M M^T is not safely positive definite (dependent landmark rows that
// are not exact duplicates — duplicates are folded exactly by the host before this point).
//
// One-sided (Hestenes) Jacobi on the rows of A = G with the rotations accumulated in V^T:
// each step rotates k/2 disjoint row pairs (round-robin tournament ordering), one warp per pair,
// rows are contiguous so every access is coalesced.  At convergence the rows of A are mutually
// orthogonal, a_i = lambda_i v_i^T, so lambda_i = |a_i| and
//     pinv(G) B = V diag(1/lambda_i, or 0 where lambda_i <= cut * lambda_max) V^T B.
// Because G squares the singular values of M, eigenvalues are resolved to ~k eps lambda_max: the
// cut is max(rcond^2, 4 k eps), i.e. singular values of M below ~sqrt(4 k eps) sigma_max are
// treated as zero (numpy keeps them down to 1e-15 sigma_max, where its answer is dominated by
// rounding noise amplified by 1/sigma).
#include "../../include/munk_b200.h"
#include "common.cuh"
#include "dgemm.cuh"
#include "workspace.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

namespace munk {
namespace {

// pair i of step s in a round-robin tournament of kk (even) players
__device__ __forceinline__ void tournament_pair(int kk, int s, int i, int& p, int& q) {
  if (i == 0) {
    p = kk - 1;
    q = s;
  } else {
    p = (s + i) % (kk - 1);
    q = (s - i + (kk - 1)) % (kk - 1);
  }
}

// One Jacobi step: rotate rows (p, q) of A and of VT for all k/2 pairs; *off = max |cos angle|.
__global__ void __launch_bounds__(256)
jacobi_step_kernel(int k, int kk, int step, double* __restrict__ A, long lda, double* __restrict__ VT,
                   long ldv, double* __restrict__ off) {
  const int pair = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pair >= kk / 2) return;
  const int lane = threadIdx.x & 31;
  int p, q;
  tournament_pair(kk, step, pair, p, q);
  if (p >= k || q >= k) return;                  // padding player (odd k)
  double* ap = A + (long)p * lda;
  double* aq = A + (long)q * lda;
  double alpha = 0.0, beta = 0.0, gamma = 0.0;
  for (int c = lane; c < k; c += 32) {
    const double x = ap[c], y = aq[c];
    alpha += x * x;
    beta += y * y;
    gamma += x * y;
  }
  for (int o = 16; o > 0; o >>= 1) {
    alpha += __shfl_xor_sync(0xffffffffu, alpha, o);
    beta += __shfl_xor_sync(0xffffffffu, beta, o);
    gamma += __shfl_xor_sync(0xffffffffu, gamma, o);
  }
  const double denom = sqrt(alpha * beta);
  if (!(denom > 0.0) || gamma == 0.0) return;
  const double rel = fabs(gamma) / denom;
  if (lane == 0) atomicMax(reinterpret_cast<unsigned long long*>(off), (unsigned long long)__double_as_longlong(rel));
  if (rel < 1e-16) return;
  // rotation that zeroes the inner product (Rutishauser's formulas)
  const double zeta = (beta - alpha) / (2.0 * gamma);
  const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
  for (int c = lane; c < k; c += 32) {
    const double x = ap[c], y = aq[c];
    ap[c] = cs * x - sn * y;
    aq[c] = sn * x + cs * y;
  }
  double* vp = VT + (long)p * ldv;
  double* vq = VT + (long)q * ldv;
  for (int c = lane; c < k; c += 32) {
    const double x = vp[c], y = vq[c];
    vp[c] = cs * x - sn * y;
    vq[c] = sn * x + cs * y;
  }
}

__global__ void identity_kernel(int k, double* __restrict__ VT, long ldv) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (c < ldv) VT[(long)r * ldv + c] = (c == r) ? 1.0 : 0.0;
}

// lambda[i] = |row i of A|
__global__ void row_norm_kernel(int k, const double* __restrict__ A, long lda, double* __restrict__ lambda) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= k) return;
  const int lane = threadIdx.x & 31;
  double s = 0.0;
  for (int c = lane; c < k; c += 32) { const double x = A[(long)r * lda + c]; s += x * x; }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) lambda[r] = sqrt(s);
}

// T[i][:] *= (lambda[i] > cut ? 1 / lambda[i] : 0)
__global__ void scale_rows_kernel(int k, int n, double* __restrict__ T, long ldt,
                                  const double* __restrict__ lambda, double cut) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (c < n) {
    const double l = lambda[r];
    T[(long)r * ldt + c] = (l > cut) ? T[(long)r * ldt + c] / l : 0.0;
  }
}

}  // namespace
}  // namespace munk

using namespace munk;

extern "C" int munk_sym_pinv_solve(munk_workspace_t* wsp, int k, double* G, long ldg, int n,
                                   const double* B, long ldb, double* Z, long ldz, double* work,
                                   double rcond, int* rank_host, void* stream_) {
  Workspace* ws = reinterpret_cast<Workspace*>(wsp);
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!ws) return -1;
  if (k <= 0 || n <= 0) { if (rank_host) *rank_host = 0; return MUNK_OK; }
  // work: VT (k x ldv) | T (k x ldt) | lambda (k) | off (1)
  const long ldv = round_up(k, 16), ldt = round_up(n, 16);
  double* VT = work;
  double* T = VT + (long)k * ldv;
  double* lambda = T + (long)k * ldt;
  double* off = lambda + round_up(k, 2);
  const int kk = k + (k & 1);

  identity_kernel<<<dim3((unsigned)((ldv + 255) / 256), k), 256, 0, stream>>>(k, VT, ldv);
  MUNK_CUDA_TRY(cudaGetLastError());
  double off_host = 1.0;
  int sweeps = 0;
  const double tol = 8.0 * 2.220446049250313e-16;
  for (; sweeps < 40 && off_host > tol; ++sweeps) {
    MUNK_CUDA_TRY(cudaMemsetAsync(off, 0, sizeof(double), stream));
    for (int s = 0; s < kk - 1; ++s) {
      jacobi_step_kernel<<<(kk / 2 + 7) / 8, 256, 0, stream>>>(k, kk, s, G, ldg, VT, ldv, off);
    }
    MUNK_CUDA_TRY(cudaGetLastError());
    MUNK_CUDA_TRY(cudaMemcpyAsync(&off_host, off, sizeof(double), cudaMemcpyDeviceToHost, stream));
    MUNK_CUDA_TRY(cudaStreamSynchronize(stream));
    ws->launches += kk - 1;
  }
  row_norm_kernel<<<(k + 7) / 8, 256, 0, stream>>>(k, G, ldg, lambda);
  MUNK_CUDA_TRY(cudaGetLastError());
  std::vector<double> lam(k);
  MUNK_CUDA_TRY(cudaMemcpyAsync(lam.data(), lambda, (size_t)k * sizeof(double), cudaMemcpyDeviceToHost, stream));
  MUNK_CUDA_TRY(cudaStreamSynchronize(stream));
  const double lmax = *std::max_element(lam.begin(), lam.end());
  const double cut = std::max(rcond * rcond, 4.0 * k * 2.220446049250313e-16) * lmax;
  int rank = 0;
  for (double l : lam) rank += (l > cut);
  if (rank_host) *rank_host = rank;

  // T = V^T B ; scale rows ; Z = V T
  GemmDesc g{};
  g.M = k; g.N = n; g.K = k;
  g.A = VT; g.lda = ldv; g.a_layout = OP_KC;
  g.B = B; g.ldb = ldb; g.b_layout = OP_MC;
  g.C = T; g.ldc = ldt; g.alpha = 1.0; g.beta = 0.0; g.flags = 0;
  MUNK_TRY(gemm_launch(ws, g, stream));
  scale_rows_kernel<<<dim3((n + 255) / 256, k), 256, 0, stream>>>(k, n, T, ldt, lambda, cut);
  MUNK_CUDA_TRY(cudaGetLastError());
  GemmDesc h{};
  h.M = k; h.N = n; h.K = k;
  h.A = VT; h.lda = ldv; h.a_layout = OP_MC;
  h.B = T; h.ldb = ldt; h.b_layout = OP_MC;
  h.C = Z; h.ldc = ldz; h.alpha = 1.0; h.beta = 0.0; h.flags = 0;
  MUNK_TRY(gemm_launch(ws, h, stream));
  ws->launches += 3;
  return MUNK_OK;
}

extern "C" size_t munk_sym_pinv_work_bytes(int k, int n) {
  const long ldv = round_up(k, 16), ldt = round_up(n, 16);
  return (size_t)((long)k * ldv + (long)k * ldt + round_up(k, 2) + 2) * sizeof(double);
}

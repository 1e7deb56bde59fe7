// Rank-revealing pseudo-inverse solve on the landmark rows themselves:
//     P = pinv(M) B,   M (k x m: k landmark rows of the source factor, m = n1),   B (k x n)
// the robust path behind np.linalg.pinv(source_C[source_idxs,:]) (munk/munk/munk.py:109), with
// numpy's semantics: singular values <= rcond * sigma_max (rcond = 1e-15) are dropped.
//
// One-sided (Hestenes) Jacobi applied from the left: plane rotations of row pairs of M until all
// rows are mutually orthogonal, R = J M with J orthogonal (k x k, accumulated).  Then
// R R^T = diag(sigma_i^2), M = J^T R is an SVD in disguise and
//     pinv(M) B = R^T diag(1 / sigma_i^2 where sigma_i > rcond sigma_max, else 0) (J B).
// The rotations act on M, not on the Gram matrix M M^T, so singular values are resolved down to
// ~eps sigma_max instead of ~sqrt(eps) sigma_max, and the final product uses the rotated rows R
// (norm sigma_i), never M^T times a 1/sigma^2-sized vector.  The caller forms P = R^T Z with the
// DMMA GEMM (munk_dgemm, A operand in MC layout).
//
// Work per sweep: k(k-1)/2 pairs x (3 dot products + 1 rotation) over m elements; one CTA per pair,
// k/2 disjoint pairs per launch (round-robin tournament), rows are contiguous (coalesced) and the
// whole k x m block (64 MB at k = 400, m = 20 000) stays in the 126 MB L2.  ~10 ms per sweep at
// that size, 6-10 sweeps: a fallback, not the hot path - the hot path (distinct rows of a kernel
// factor, condition <= ~3) is the Gram/Cholesky solve in device.py, whose applicability is decided
// on the device by a rigorous condition bound.
#include "../../include/munk_b200.h"
#include "common.cuh"
#include "dgemm.cuh"
#include "workspace.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

namespace munk {
namespace {

constexpr int JT = 256;    // threads per row pair

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

__device__ __forceinline__ double block_sum(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < JT / 32; ++w) s += sh[w];
  return s;
}

// One Jacobi step: for every pair (p, q) of this tournament round rotate rows p, q of R (length
// m) and of J (length k) so that the two rows of R become orthogonal.
//   *off   = max over pairs of |<r_p, r_q>| / (|r_p| |r_q|)   (convergence measure)
//   floor2 = rows whose squared norm is below it are numerically zero (far below the rank cut):
//            they are not rotated against (their direction is rounding noise)
__global__ void __launch_bounds__(JT)
rows_jacobi_step_kernel(int k, int kk, int step, int m, double* __restrict__ R, long ldr,
                        double* __restrict__ J, long ldj, double floor2, double* __restrict__ off) {
  __shared__ double sh[JT / 32];
  __shared__ double rot[2];
  int p, q;
  tournament_pair(kk, step, blockIdx.x, p, q);
  if (p >= k || q >= k) return;                  // padding player (odd k)
  double* rp = R + (long)p * ldr;
  double* rq = R + (long)q * ldr;
  double alpha = 0.0, beta = 0.0, gamma = 0.0;
  for (int c = threadIdx.x; c < m; c += JT) {
    const double x = rp[c], y = rq[c];
    alpha = fma(x, x, alpha);
    beta = fma(y, y, beta);
    gamma = fma(x, y, gamma);
  }
  alpha = block_sum(alpha, sh);
  beta = block_sum(beta, sh);
  gamma = block_sum(gamma, sh);
  if (threadIdx.x == 0) {
    double cs = 1.0, sn = 0.0;
    const double denom = sqrt(alpha) * sqrt(beta);
    if (denom > 0.0 && gamma != 0.0 && alpha > floor2 && beta > floor2) {
      const double rel = fabs(gamma) / denom;
      atomicMax(reinterpret_cast<unsigned long long*>(off), (unsigned long long)__double_as_longlong(rel));
      if (rel >= 1e-16) {
        // rotation that zeroes the inner product (Rutishauser's formulas)
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        cs = 1.0 / sqrt(1.0 + t * t);
        sn = cs * t;
      }
    }
    rot[0] = cs;
    rot[1] = sn;
  }
  __syncthreads();
  const double cs = rot[0], sn = rot[1];
  if (sn == 0.0) return;
  for (int c = threadIdx.x; c < m; c += JT) {
    const double x = rp[c], y = rq[c];
    rp[c] = cs * x - sn * y;
    rq[c] = sn * x + cs * y;
  }
  double* jp = J + (long)p * ldj;
  double* jq = J + (long)q * ldj;
  for (int c = threadIdx.x; c < k; c += JT) {
    const double x = jp[c], y = jq[c];
    jp[c] = cs * x - sn * y;
    jq[c] = sn * x + cs * y;
  }
}

__global__ void identity_kernel(int k, double* __restrict__ J, long ldj) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (c < ldj) J[(long)r * ldj + c] = (c == r) ? 1.0 : 0.0;
}

// sigma[i] = |row i of R|
__global__ void __launch_bounds__(JT)
row_norm_kernel(int m, const double* __restrict__ R, long ldr, double* __restrict__ sigma) {
  __shared__ double sh[JT / 32];
  const double* r = R + (long)blockIdx.x * ldr;
  double s = 0.0;
  for (int c = threadIdx.x; c < m; c += JT) s = fma(r[c], r[c], s);
  s = block_sum(s, sh);
  if (threadIdx.x == 0) sigma[blockIdx.x] = sqrt(s);
}

// T[i][:] *= (sigma[i] > cut ? 1 / sigma[i]^2 : 0)
__global__ void scale_rows_kernel(int n, double* __restrict__ T, long ldt,
                                  const double* __restrict__ sigma, double cut) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (c < n) {
    const double s = sigma[r];
    T[(long)r * ldt + c] = (s > cut) ? (T[(long)r * ldt + c] / s) / s : 0.0;
  }
}

}  // namespace
}  // namespace munk

using namespace munk;

extern "C" int munk_pinv_rows_solve(munk_workspace_t* wsp, int k, int m, double* R, long ldr, int n,
                                    const double* B, long ldb, double* Z, long ldz, double* work,
                                    double rcond, int* rank_host, double* sigma_host, void* stream_) {
  Workspace* ws = reinterpret_cast<Workspace*>(wsp);
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!ws) return -1;
  if (k < 0) return -2;
  if (m < 0) return -3;
  if (ldr < m) return -5;
  if (n < 0) return -6;
  if (rank_host) *rank_host = 0;
  if (k == 0 || m == 0) return MUNK_OK;
  DeviceGuard guard(ws->device);
  // work: J (k x ldj) | sigma (k) | off (2)
  const long ldj = round_up(k, 16);
  double* J = work;
  double* sigma = J + (long)k * ldj;
  double* off = sigma + round_up(k, 2);
  const int kk = k + (k & 1);

  identity_kernel<<<dim3((unsigned)((ldj + 255) / 256), k), 256, 0, stream>>>(k, J, ldj);
  MUNK_CUDA_TRY(cudaGetLastError());
  std::vector<double> sg(k);
  auto norms = [&]() -> int {
    row_norm_kernel<<<k, JT, 0, stream>>>(m, R, ldr, sigma);
    MUNK_CUDA_TRY(cudaGetLastError());
    MUNK_CUDA_TRY(cudaMemcpyAsync(sg.data(), sigma, (size_t)k * sizeof(double), cudaMemcpyDeviceToHost, stream));
    MUNK_CUDA_TRY(cudaStreamSynchronize(stream));
    return MUNK_OK;
  };
  MUNK_TRY(norms());
  double off_host = 1.0;
  const double tol = 8.0 * 2.220446049250313e-16;
  for (int sweep = 0; sweep < 60 && off_host > tol && kk > 1; ++sweep) {
    const double big = *std::max_element(sg.begin(), sg.end());
    const double floor2 = (1e-17 * big) * (1e-17 * big);
    MUNK_CUDA_TRY(cudaMemsetAsync(off, 0, sizeof(double), stream));
    for (int s = 0; s < kk - 1; ++s)
      rows_jacobi_step_kernel<<<kk / 2, JT, 0, stream>>>(k, kk, s, m, R, ldr, J, ldj, floor2, off);
    MUNK_CUDA_TRY(cudaGetLastError());
    MUNK_CUDA_TRY(cudaMemcpyAsync(&off_host, off, sizeof(double), cudaMemcpyDeviceToHost, stream));
    MUNK_TRY(norms());
    ws->launches += kk;
  }
  const double smax = *std::max_element(sg.begin(), sg.end());
  const double cut = rcond * smax;                       // numpy: s > rcond * amax(s) is kept
  int rank = 0;
  for (double s : sg) rank += (s > cut);
  if (rank_host) *rank_host = rank;
  if (sigma_host) std::copy(sg.begin(), sg.end(), sigma_host);
  if (n == 0) return MUNK_OK;

  // Z = diag(sigma^+2) (J B)
  GemmDesc g{};
  g.M = k; g.N = n; g.K = k;
  g.A = J; g.lda = ldj; g.a_layout = OP_KC;
  g.B = B; g.ldb = ldb; g.b_layout = OP_MC;
  g.C = Z; g.ldc = ldz; g.alpha = 1.0; g.beta = 0.0; g.flags = 0;
  MUNK_TRY(gemm_launch(ws, g, stream));
  scale_rows_kernel<<<dim3((n + 255) / 256, k), 256, 0, stream>>>(n, Z, ldz, sigma, cut);
  MUNK_CUDA_TRY(cudaGetLastError());
  ws->launches += 2;
  return MUNK_OK;
}

extern "C" size_t munk_pinv_rows_work_bytes(int k) {
  const long ldj = round_up(k, 16);
  return (size_t)((long)k * ldj + round_up(k, 2) + 2) * sizeof(double);
}

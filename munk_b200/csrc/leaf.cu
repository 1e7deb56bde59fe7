// One-CTA Cholesky + inverse of a 128x128 diagonal block (the leaf of the recursion in factor.cu).
//
// The leaf sits on the critical path of the factorisation 157 times for n = 20 000, so it is
// built for latency, not throughput:
//   * the block lives in shared memory with row stride 132 doubles (= 4 mod 16), which makes
//     every DMMA fragment load below bank-conflict-free in both operand orientations;
//   * panels of 16 columns.  Each of the 8 warps holds the 16x16 diagonal block of the panel
//     redundantly in lanes 0-15 (one row per lane, 16 registers) and 16 of the rows below it in
//     lanes 16-31, and eliminates column by column with warp shuffles only (pivot broadcast,
//     L[t][j] broadcast): the diagonal factorisation and the panel solve are one pass with no
//     block-level synchronisation;
//   * the trailing update of the panel (rank 16) runs on the FP64 tensor cores (DMMA m8n8k4),
//     8x8 tiles of the lower triangle spread over the warps;
//   * the inverse is formed by recursive doubling: 16x16 diagonal blocks in registers (one per
//     warp), then three merge levels X21 = -X22 (L21 X11), all DMMA.
// A block smaller than 128 is padded with the identity, which factors and inverts to itself.
#include "common.cuh"
#include "leaf.cuh"

namespace munk {
namespace {

constexpr int NB = MUNK_LEAF;       // 128
constexpr int PW = 16;              // panel width
constexpr int LDM = 132;            // shared row stride of the block
constexpr int LDT = 68;             // shared row stride of the merge temporary (64 x 64)
constexpr int LEAF_THREADS = 256;
constexpr unsigned FULL = 0xffffffffu;

// triangular tile index t -> (tr, tc) with tc <= tr, t = tr (tr + 1) / 2 + tc
__device__ __forceinline__ void tri_tile(int t, int& tr, int& tc) {
  tr = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
  while (tr * (tr + 1) / 2 > t) --tr;
  while ((tr + 1) * (tr + 2) / 2 <= t) ++tr;
  tc = t - tr * (tr + 1) / 2;
}

__global__ void __launch_bounds__(LEAF_THREADS, 1)
potrf_leaf_kernel(double* __restrict__ A, long lda, int nb, double* __restrict__ Linv, int gidx,
                  int* __restrict__ info) {
  extern __shared__ double sm[];
  double* __restrict__ M = sm;              // [128][LDM]
  double* __restrict__ T = sm + NB * LDM;   // [64][LDT]
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int r8 = lane >> 2, j4 = lane & 3;

  // M = [ tril(A) 0 ; 0 I ]
  for (int idx = tid; idx < NB * NB; idx += LEAF_THREADS) {
    const int r = idx >> 7, c = idx & 127;
    double v = (r == c) ? 1.0 : 0.0;
    if (r < nb && c <= r) v = A[(long)r * lda + c];
    M[r * LDM + c] = v;
  }
  __syncthreads();

  int bad = 0;   // identical in every thread: pivots are warp-broadcast values of the same data
  for (int p = 0; p < NB / PW; ++p) {
    const int c0 = p * PW;
    // ---- diagonal block + panel rows, one row per lane, shuffles only -----------------------
    {
      int row = c0 + lane;
      bool active = true;
      if (lane >= 16) {
        row = c0 + 16 * warp + lane;          // = c0 + 16 + 16*warp + (lane - 16)
        active = row < NB;
      }
      const double* src = M + (active ? row : c0) * LDM + c0;
      double a[PW];
#pragma unroll
      for (int t = 0; t < PW; ++t) a[t] = src[t];
#pragma unroll
      for (int j = 0; j < PW; ++j) {
        const double d = __shfl_sync(FULL, a[j], j);
        if (!(d > 0.0) && !bad) bad = c0 + j + 1;
        // one reciprocal square root instead of sqrt + divide on the serial chain
        // (rsqrt is <= 1 ulp; L[j][j] = d * rsqrt(d) is then within 2 ulp of sqrt(d))
        const double rinv = rsqrt(d);
        a[j] = (lane == j) ? d * rinv : a[j] * rinv;
#pragma unroll
        for (int t = j + 1; t < PW; ++t) {
          const double l = __shfl_sync(FULL, a[j], t);      // L[c0+t][c0+j]
          a[t] = fma(-a[j], l, a[t]);
        }
      }
      if (lane < 16) {
        if (warp == 0) {
          double* dst = M + row * LDM + c0;
#pragma unroll
          for (int t = 0; t < PW; ++t) dst[t] = (t <= lane) ? a[t] : 0.0;
        }
      } else if (active) {
        double* dst = M + row * LDM + c0;
#pragma unroll
        for (int t = 0; t < PW; ++t) dst[t] = a[t];
      }
    }
    __syncthreads();
    // ---- trailing update A22 -= X X^T on the tensor cores (lower 8x8 tiles) ------------------
    {
      const int R0 = c0 + PW;
      const int m = (NB - R0) >> 3;
      const int ntiles = m * (m + 1) / 2;
      for (int t = warp; t < ntiles; t += LEAF_THREADS / 32) {
        int tr, tc;
        tri_tile(t, tr, tc);
        double* cp = M + (R0 + 8 * tr + r8) * LDM + R0 + 8 * tc + 2 * j4;
        double v0 = cp[0], v1 = cp[1];
        const double* ap = M + (R0 + 8 * tr + r8) * LDM + c0 + j4;
        const double* bp = M + (R0 + 8 * tc + r8) * LDM + c0 + j4;
#pragma unroll
        for (int s = 0; s < PW / 4; ++s) dmma884(v0, v1, -ap[4 * s], bp[4 * s]);
        cp[0] = v0;
        cp[1] = v1;
      }
    }
    __syncthreads();
  }
  if (bad && tid == 0) atomicCAS(info, 0, gidx + bad);   // keep the first failure

  // ---- L out (zero strict upper) ------------------------------------------------------------
  for (int idx = tid; idx < nb * NB; idx += LEAF_THREADS) {
    const int r = idx >> 7, c = idx & 127;
    if (c < nb) A[(long)r * lda + c] = (c <= r) ? M[r * LDM + c] : 0.0;
  }
  __syncthreads();

  // ---- inverse, level 0: warp w inverts the 16x16 diagonal block w (lane = column) ------------
  {
    const int o = 16 * warp;
    double x[PW];
    if (lane < 16) {
#pragma unroll
      for (int i = 0; i < PW; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < i; ++k) acc = fma(M[(o + i) * LDM + o + k], x[k], acc);   // x[k] = 0 for k < lane
        x[i] = (((i == lane) ? 1.0 : 0.0) - acc) / M[(o + i) * LDM + o + i];
      }
    }
    __syncwarp();
    if (lane < 16) {
#pragma unroll
      for (int i = 0; i < PW; ++i) M[(o + i) * LDM + o + lane] = x[i];
    }
  }
  __syncthreads();

  // ---- merge levels: X21 = -X22 (L21 X11) for block sizes 16, 32, 64 ---------------------------
  for (int b = 16; b < NB; b <<= 1) {
    const int tb = b >> 3;                  // 8x8 tiles per block side
    const int tpb = tb * tb;
    const int ntiles = (NB / (2 * b)) * tpb;
    // T_q = L21 X11   (X11 lower: k >= column)
    for (int t = warp; t < ntiles; t += LEAF_THREADS / 32) {
      const int q = t / tpb, tt = t % tpb, tr = tt / tb, tc = tt % tb;
      const int o = 2 * b * q;
      double v0 = 0.0, v1 = 0.0;
      const double* ap = M + (o + b + 8 * tr + r8) * LDM + o + j4;
      const double* bp = M + (o + j4) * LDM + o + 8 * tc + r8;
      for (int k = 8 * tc; k < b; k += 4) dmma884(v0, v1, ap[k], bp[k * LDM]);
      double* d = T + (8 * tr + r8) * LDT + q * b + 8 * tc + 2 * j4;
      d[0] = v0;
      d[1] = v1;
    }
    __syncthreads();
    // X21 = -X22 T_q  (X22 lower: k <= row)
    for (int t = warp; t < ntiles; t += LEAF_THREADS / 32) {
      const int q = t / tpb, tt = t % tpb, tr = tt / tb, tc = tt % tb;
      const int o = 2 * b * q;
      double v0 = 0.0, v1 = 0.0;
      const double* ap = M + (o + b + 8 * tr + r8) * LDM + o + b + j4;
      const double* bp = T + j4 * LDT + q * b + 8 * tc + r8;
      for (int k = 0; k < 8 * tr + 8; k += 4) dmma884(v0, v1, ap[k], bp[k * LDT]);
      double* d = M + (o + b + 8 * tr + r8) * LDM + o + 8 * tc + 2 * j4;
      d[0] = -v0;
      d[1] = -v1;
    }
    __syncthreads();
  }

  // ---- X out: row-major 128 x 128, lower, zero outside the nb x nb block -----------------------
  for (int idx = tid; idx < NB * NB; idx += LEAF_THREADS) {
    const int r = idx >> 7, c = idx & 127;
    Linv[idx] = (r < nb && c <= r) ? M[r * LDM + c] : 0.0;
  }
}

}  // namespace

int launch_potrf_leaf(double* A, long lda, int nb, double* Linv, int gidx, int* info,
                      cudaStream_t stream) {
  static bool configured[MUNK_MAX_DEVICES] = {};       // the attribute is per device
  const int smem = (NB * LDM + 64 * LDT) * 8;
  int dev = 0;
  MUNK_CUDA_TRY(cudaGetDevice(&dev));
  if (dev >= MUNK_MAX_DEVICES || !configured[dev]) {
    MUNK_CUDA_TRY(cudaFuncSetAttribute(potrf_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (dev < MUNK_MAX_DEVICES) configured[dev] = true;
  }
  potrf_leaf_kernel<<<1, LEAF_THREADS, smem, stream>>>(A, lda, nb, Linv, gidx, info);
  MUNK_CUDA_TRY(cudaGetLastError());
  return MUNK_OK;
}

}  // namespace munk

// Device-side body of the 128 x 128 Cholesky leaf (shared by potrf_leaf_kernel in leaf.cu and the
// fused diagonal-block kernel in diag.cu): the block lives in shared memory M[128][LDM]; see leaf.cu
// for the algorithm.
#pragma once
#include "common.cuh"

namespace munk {
namespace leafdev {

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


// In-place Cholesky of the lower triangle of M (block padded with the identity).  Returns 0 or
// 1 + index of the first non-positive pivot (the same value in every thread).  Needs 256 threads.
__device__ __forceinline__ int leaf_factor(double* __restrict__ M) {
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int r8 = lane >> 2, j4 = lane & 3;
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
  return bad;
}

// M (lower triangular factor) <- its inverse, in place; T is a [64][LDT] temporary.
__device__ __forceinline__ void leaf_invert(double* __restrict__ M, double* __restrict__ T) {
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int r8 = lane >> 2, j4 = lane & 3;
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

}

}  // namespace leafdev
}  // namespace munk

// Panel solve in ONE launch:  B (m x s, in place) <- B L^-T  for an s x s lower-triangular block
// L (s <= 512) whose 128 x 128 diagonal blocks have their inverses stored (the leaf inverses the
// Cholesky leaf kernel leaves behind).
//
// The rows of a right-side triangular solve are independent, so a CTA takes a strip of 32 rows and
// runs the whole blocked substitution on it with no inter-CTA dependency:
//     for k = 0 .. s/128-1:   Y_k = B_k - sum_{j<k} X_j L_kj^T ;   X_k = Y_k inv(L_kk)^T
// The strip (32 x s) lives in shared memory, the blocks of L and the leaf inverses stream from
// L2 (the block is at most 2 MB); all products run on the FP64 tensor cores (DMMA m8n8k4), warp w
// owning columns [16 w, 16 w + 16) of the current 128-wide block.
//
// This replaces the recursive solve (trsm_rec: ~14 dependent small GEMM launches with split-K
// reductions for s = 512) where latency matters: the first panel rows after a diagonal block is
// factored (task TS_N of the block-column driver, on the critical chain of the factorisation; a
// chain kernel that is launched waits until an SM drains, every dependent launch again).
// Reference call replaced: the panel solve inside np.linalg.inv / scipy eigh of
// munk/munk/munk.py:83,87 (LAPACK dtrsm).
#include "common.cuh"
#include "leaf.cuh"

namespace munk {
namespace {

constexpr int SR = 32;              // rows per CTA
constexpr int SBLK = MUNK_LEAF;     // 128
constexpr int SMAX = 512;           // largest s
constexpr int SLD = SMAX + 4;       // shared row stride (= 4 mod 16 doubles: conflict-free DMMA fragments)
constexpr int STHREADS = 256;

// acc[rt][ct] (+)= A (rows 8 rt .. +8 of the strip, columns a0 .. a0+tn of shared X)
//                 * Bm^T, Bm = rows c0 + 8 ct .. +8 (global, row stride ldb, first column b0) x tn
// t runs over [0, tn); rows of Bm at or beyond `bm_rows` are treated as zero.
__device__ __forceinline__ void strip_mma(double (&acc)[4][2][2], const double* __restrict__ Xs, int a0,
                                          const double* __restrict__ Bm, long ldb, int c0, int bm_rows,
                                          int tn, int lane) {
  const int r8 = lane >> 2, j4 = lane & 3;
  const double* b0p = Bm + (long)(c0 + r8) * ldb + j4;
  const double* b1p = b0p + 8 * ldb;
  const bool ok0 = c0 + r8 < bm_rows, ok1 = c0 + 8 + r8 < bm_rows;
  const double* ap = Xs + r8 * SLD + a0 + j4;
#pragma unroll 8
  for (int t = 0; t < tn; t += 4) {
    const double b0 = ok0 ? __ldg(b0p + t) : 0.0;
    const double b1 = ok1 ? __ldg(b1p + t) : 0.0;
#pragma unroll
    for (int rt = 0; rt < 4; ++rt) {
      const double a = ap[rt * 8 * SLD + t];
      dmma884(acc[rt][0][0], acc[rt][0][1], a, b0);
      dmma884(acc[rt][1][0], acc[rt][1][1], a, b1);
    }
  }
}

__global__ void __launch_bounds__(STHREADS, 1)
trsm_strip_kernel(int m, int s, double* __restrict__ B, long ldb, const double* __restrict__ L, long ldl,
                  const double* __restrict__ linv) {
  extern __shared__ double Xs[];            // [SR][SLD]
  const int row0 = blockIdx.x * SR;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r8 = lane >> 2, j4 = lane & 3;
  const int nbk = (s + SBLK - 1) / SBLK;
  const int sp = nbk * SBLK;                // s padded to whole blocks

  // strip in (zero beyond m rows / s columns)
  for (int idx = tid; idx < SR * sp; idx += STHREADS) {
    const int r = idx / sp, c = idx - r * sp;
    Xs[r * SLD + c] = (row0 + r < m && c < s) ? B[(long)(row0 + r) * ldb + c] : 0.0;
  }
  __syncthreads();

  const int c0 = 16 * warp;                 // this warp's columns inside the current block
  for (int k = 0; k < nbk; ++k) {
    const int kb = k * SBLK;
    const int rows_k = min(SBLK, s - kb);   // valid rows of block row k of L
    double acc[4][2][2];
#pragma unroll
    for (int rt = 0; rt < 4; ++rt)
#pragma unroll
      for (int ct = 0; ct < 2; ++ct) acc[rt][ct][0] = acc[rt][ct][1] = 0.0;
    // acc = sum_{j<k} X_j L_kj^T   (one pass over columns [0, kb) of block row k of L)
    if (k > 0) strip_mma(acc, Xs, 0, L + (long)kb * ldl, ldl, c0, rows_k, kb, lane);
    // Y_k = B_k - acc, back into the strip
#pragma unroll
    for (int rt = 0; rt < 4; ++rt)
#pragma unroll
      for (int ct = 0; ct < 2; ++ct) {
        double* p = Xs + (8 * rt + r8) * SLD + kb + c0 + 8 * ct + 2 * j4;
        p[0] -= acc[rt][ct][0];
        p[1] -= acc[rt][ct][1];
        acc[rt][ct][0] = acc[rt][ct][1] = 0.0;
      }
    __syncthreads();
    // X_k = Y_k inv(L_kk)^T ; the inverse is lower triangular: only t <= column contributes
    strip_mma(acc, Xs, kb, linv + (long)k * SBLK * SBLK, SBLK, c0, SBLK, c0 + 16, lane);
    __syncthreads();                        // every warp has read Y_k
#pragma unroll
    for (int rt = 0; rt < 4; ++rt)
#pragma unroll
      for (int ct = 0; ct < 2; ++ct) {
        double* p = Xs + (8 * rt + r8) * SLD + kb + c0 + 8 * ct + 2 * j4;
        p[0] = acc[rt][ct][0];
        p[1] = acc[rt][ct][1];
      }
    __syncthreads();
  }

  for (int idx = tid; idx < SR * sp; idx += STHREADS) {
    const int r = idx / sp, c = idx - r * sp;
    if (row0 + r < m && c < s) B[(long)(row0 + r) * ldb + c] = Xs[r * SLD + c];
  }
}

}  // namespace

int launch_trsm_strip(double* B, long ldb, int m, const double* L, long ldl, int s, const double* linv,
                      cudaStream_t stream) {
  if (m <= 0 || s <= 0) return MUNK_OK;
  if (s > SMAX) {
    set_last_error("trsm_strip: block of %d columns (max %d)", s, SMAX);
    return -6;
  }
  static bool configured[MUNK_MAX_DEVICES] = {};
  const int smem = SR * SLD * 8;
  int dev = 0;
  MUNK_CUDA_TRY(cudaGetDevice(&dev));
  if (dev >= MUNK_MAX_DEVICES || !configured[dev]) {
    MUNK_CUDA_TRY(cudaFuncSetAttribute(trsm_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (dev < MUNK_MAX_DEVICES) configured[dev] = true;
  }
  trsm_strip_kernel<<<(m + SR - 1) / SR, STHREADS, smem, stream>>>(m, s, B, ldb, L, ldl, linv);
  MUNK_CUDA_TRY(cudaGetLastError());
  return MUNK_OK;
}

}  // namespace munk

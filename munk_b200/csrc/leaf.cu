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
#include "leaf_device.cuh"

namespace munk {
namespace {

using namespace leafdev;

__global__ void __launch_bounds__(LEAF_THREADS, 1)
potrf_leaf_kernel(double* __restrict__ A, long lda, int nb, double* __restrict__ Linv, int gidx,
                  int* __restrict__ info) {
  extern __shared__ double sm[];
  double* __restrict__ M = sm;              // [128][LDM]
  double* __restrict__ T = sm + NB * LDM;   // [64][LDT]
  const int tid = threadIdx.x;

  // M = [ tril(A) 0 ; 0 I ]
  for (int idx = tid; idx < NB * NB; idx += LEAF_THREADS) {
    const int r = idx >> 7, c = idx & 127;
    double v = (r == c) ? 1.0 : 0.0;
    if (r < nb && c <= r) v = A[(long)r * lda + c];
    M[r * LDM + c] = v;
  }
  __syncthreads();

  const int bad = leaf_factor(M);
  if (bad && tid == 0) atomicCAS(info, 0, gidx + bad);   // keep the first failure

  // ---- L out (zero strict upper) ------------------------------------------------------------
  for (int idx = tid; idx < nb * NB; idx += LEAF_THREADS) {
    const int r = idx >> 7, c = idx & 127;
    if (c < nb) A[(long)r * lda + c] = (c <= r) ? M[r * LDM + c] : 0.0;
  }
  __syncthreads();

  leaf_invert(M, T);

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

// Cholesky of a diagonal block of 256 / 384 / 512 columns in ONE launch.
//
// The diagonal block F(c) sits on the only serial chain of the factorisation
// (LA_D -> F -> TS_N, factor.cu / dist.cu).  As a recursion it is ~30 dependent small launches
// (4 leaves, triangular solves and updates with split-K reductions), ~0.32 ms on an idle GPU - but
// the GPU is not idle: every SM holds one 85 us trailing-update tile, each dependent launch waits for
// an SM to drain, and the block takes 0.6 ms (one GPU, mid-matrix) to 1.5 ms and more (8 GPUs,
// profiles/r02_dist_trace_n8.log).  Here the whole block is one grid of NCTA co-operating CTAs that
// waits for SMs once and then runs the right-looking algorithm on 128-column steps with grid-wide
// barriers in global memory (atomic counter, L2-coherent loads):
//     step k:  CTA 0: leaf Cholesky + inverse of A_kk (shared memory, warp shuffles, DMMA: leaf.cu)
//              barrier
//              all  : panel  A_ik <- A_ik inv(L_kk)^T  for i > k, 64-row strips (DMMA)
//              barrier
//              all  : update A_ij <- A_ij - A_ik A_jk^T for k < j <= i, 64 x 64 tiles (DMMA)
//              barrier
// Outputs are those of the recursion: L in place (zeros above the diagonal inside the 128 x 128
// diagonal blocks), the leaf inverses in `linv`, the non-SPD flag in *info.
// Reference call replaced: the diagonal-block factorisation inside np.linalg.inv / scipy eigh of
// munk/munk/munk.py:83,87 (LAPACK dpotf2 / dpotrf).
#include "common.cuh"
#include "leaf.cuh"
#include "leaf_device.cuh"

namespace munk {
namespace {

using namespace leafdev;

constexpr int DN = 24;                 // CTAs of the grid (>= the 21 update tiles of the first step)
constexpr int DLD = NB + 4;            // shared row stride of staged operands (= 4 mod 16)
constexpr int DSMEM = (64 * DLD + NB * DLD) * 8;   // A strip (64 rows) + B operand (128 rows): 202 752 B

// Every CTA of the grid arrives; returns when all have (release / acquire through L2).
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    while (*reinterpret_cast<volatile unsigned*>(counter) < target) __nanosleep(40);
    __threadfence();
  }
  __syncthreads();
}

// dst[r][0:cols) (shared, stride DLD) <- src[r][0:cols) (global, stride ld), r < rows; L2 loads.
__device__ __forceinline__ void stage(double* __restrict__ dst, const double* __restrict__ src, long ld,
                                      int rows, int cols) {
  const int pairs = cols >> 1;
  for (int idx = threadIdx.x; idx < rows * pairs; idx += LEAF_THREADS) {
    const int r = idx / pairs, c = (idx - r * pairs) * 2;
    const double2 v = __ldcg(reinterpret_cast<const double2*>(src + (long)r * ld + c));
    dst[r * DLD + c] = v.x;
    dst[r * DLD + c + 1] = v.y;
  }
}

// acc[ct] (8 x 8 tile: rows 8 warp .. +8 of As, columns 8 ct .. +8) = As (64 x K) Bs (8 NCT x K)^T
template <int NCT>
__device__ __forceinline__ void mma_nt(double (&acc)[NCT][2], const double* __restrict__ As,
                                       const double* __restrict__ Bs, int K) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r8 = lane >> 2, j4 = lane & 3;
#pragma unroll
  for (int ct = 0; ct < NCT; ++ct) acc[ct][0] = acc[ct][1] = 0.0;
  const double* ap = As + (8 * warp + r8) * DLD + j4;
  const double* bp = Bs + r8 * DLD + j4;
  for (int t = 0; t < K; t += 4) {
    const double a = ap[t];
#pragma unroll
    for (int ct = 0; ct < NCT; ++ct) dmma884(acc[ct][0], acc[ct][1], a, bp[ct * 8 * DLD + t]);
  }
}

__global__ void __launch_bounds__(LEAF_THREADS, 1)
potrf_diag_kernel(double* __restrict__ A, long lda, int q, double* __restrict__ linv, int gidx,
                  int* __restrict__ info, unsigned* __restrict__ counter) {
  extern __shared__ double sm[];
  double* __restrict__ M = sm;                 // leaf: [128][LDM] + T [64][LDT];  LDM == DLD
  double* __restrict__ T = sm + NB * LDM;
  double* __restrict__ As = sm;                // panel / update: A operand [64][DLD]
  double* __restrict__ Bs = sm + 64 * DLD;     //                 B operand [128][DLD]
  const int tid = threadIdx.x, cta = blockIdx.x;
  const int warp = tid >> 5, lane = tid & 31, r8 = lane >> 2, j4 = lane & 3;
  unsigned epoch = 0;

  for (int k = 0; k < q; ++k) {
    const int kb = k * NB;
    double* Akk = A + (long)kb * lda + kb;
    double* Xkk = linv + (long)k * NB * NB;
    // ---- leaf: Cholesky + inverse of the diagonal block (CTA 0) ------------------------------
    if (cta == 0) {
      for (int idx = tid; idx < NB * NB; idx += LEAF_THREADS) {
        const int r = idx >> 7, c = idx & 127;
        M[r * LDM + c] = (c <= r) ? __ldcg(Akk + (long)r * lda + c) : 0.0;
      }
      __syncthreads();
      const int bad = leaf_factor(M);
      if (bad && tid == 0) atomicCAS(info, 0, gidx + kb + bad);
      for (int idx = tid; idx < NB * NB; idx += LEAF_THREADS) {
        const int r = idx >> 7, c = idx & 127;
        Akk[(long)r * lda + c] = (c <= r) ? M[r * LDM + c] : 0.0;
      }
      __syncthreads();
      leaf_invert(M, T);
      for (int idx = tid; idx < NB * NB; idx += LEAF_THREADS) {
        const int r = idx >> 7, c = idx & 127;
        Xkk[idx] = (c <= r) ? M[r * LDM + c] : 0.0;
      }
    }
    if (k + 1 == q) break;
    grid_barrier(counter, (++epoch) * DN);
    // ---- panel: A_ik <- A_ik inv(L_kk)^T, 64-row strips of the rows below ----------------------
    const int nstrips = (q - 1 - k) * 2;
    for (int t = cta; t < nstrips; t += DN) {
      double* Ai = A + (long)(kb + NB + 64 * t) * lda + kb;
      stage(As, Ai, lda, 64, NB);
      stage(Bs, Xkk, NB, NB, NB);
      __syncthreads();
      double acc[16][2];
      mma_nt<16>(acc, As, Bs, NB);
#pragma unroll
      for (int ct = 0; ct < 16; ++ct) {
        double* p = Ai + (long)(8 * warp + r8) * lda + 8 * ct + 2 * j4;
        *reinterpret_cast<double2*>(p) = make_double2(acc[ct][0], acc[ct][1]);
      }
      __syncthreads();
    }
    grid_barrier(counter, (++epoch) * DN);
    // ---- update: A_ij <- A_ij - A_ik A_jk^T, 64 x 64 tiles touching the lower triangle ---------
    // tile (ti, tj) in units of 64 rows / columns of the trailing block, tj <= ti
    const int h = (q - 1 - k) * 2;              // trailing block is h x h tiles
    const int ntiles = h * (h + 1) / 2;
    for (int t = cta; t < ntiles; t += DN) {
      int ti, tj;
      tri_tile(t, ti, tj);
      const double* Xi = A + (long)(kb + NB + 64 * ti) * lda + kb;
      const double* Xj = A + (long)(kb + NB + 64 * tj) * lda + kb;
      stage(As, Xi, lda, 64, NB);
      stage(Bs, Xj, lda, 64, NB);
      __syncthreads();
      double acc[8][2];
      mma_nt<8>(acc, As, Bs, NB);
      double* Cij = A + (long)(kb + NB + 64 * ti) * lda + kb + NB + 64 * tj;
#pragma unroll
      for (int ct = 0; ct < 8; ++ct) {
        double2* p = reinterpret_cast<double2*>(Cij + (long)(8 * warp + r8) * lda + 8 * ct + 2 * j4);
        double2 c = __ldcg(p);
        c.x -= acc[ct][0];
        c.y -= acc[ct][1];
        *p = c;
      }
      __syncthreads();
    }
    grid_barrier(counter, (++epoch) * DN);
  }
}

}  // namespace

// A (w x w, w = 128 q, q = 2..4): in-place lower Cholesky; leaf inverses to linv; *info as the leaf.
// `counter` is a device word used by the grid barrier; it is zeroed here on `stream`.
int launch_potrf_diag(double* A, long lda, int w, double* linv, int gidx, int* info, unsigned* counter,
                      cudaStream_t stream) {
  if (w % NB != 0 || w < 2 * NB || w > 4 * NB || (lda & 1) || ((uintptr_t)A & 15)) {
    set_last_error("potrf_diag: unsupported block (w=%d, lda=%ld)", w, lda);
    return -3;
  }
  static_assert(LDM == DLD, "the leaf and the staged operands share one shared-memory layout");
  static bool configured[MUNK_MAX_DEVICES] = {};
  int dev = 0;
  MUNK_CUDA_TRY(cudaGetDevice(&dev));
  if (dev >= MUNK_MAX_DEVICES || !configured[dev]) {
    MUNK_CUDA_TRY(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DSMEM));
    if (dev < MUNK_MAX_DEVICES) configured[dev] = true;
  }
  MUNK_CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(unsigned), stream));
  potrf_diag_kernel<<<DN, LEAF_THREADS, DSMEM, stream>>>(A, lda, w / NB, linv, gidx, info, counter);
  MUNK_CUDA_TRY(cudaGetLastError());
  return MUNK_OK;
}

}  // namespace munk

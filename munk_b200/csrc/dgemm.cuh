// Internal interface of the FP64 tensor-core GEMM (dgemm.cu).
#pragma once
#include "common.cuh"

namespace munk {

// Operand storage.  "KC": the contraction index k is the contiguous one,
// X[i*ld + k]; "MC": the free index (m or n) is contiguous, X[k*ld + i].
enum : int { OP_KC = 0, OP_MC = 1 };

// Structure flags: which output tiles exist and which k-range each tile needs.
enum : int {
  GF_TILE_LOWER = 1,   // only tiles that intersect the lower triangle (n0 < m0 + BM)
  GF_KLO_M = 2,        // operands vanish for k <  m0      (k >= row index)
  GF_KLO_N = 4,        // operands vanish for k <  n0      (k >= col index)
  GF_KHI_M = 8,        // operands vanish for k >= m0 + BM (k <= row index)
  GF_KHI_N = 16,       // operands vanish for k >= n0 + BN (k <= col index)
  GF_SYM_MIRROR = 32,  // also store the transposed tile (symmetric result, with GF_TILE_LOWER)
};

struct GemmDesc {
  int M, N, K;
  const double* A; long lda; int a_layout;  // M x K operand
  const double* B; long ldb; int b_layout;  // N x K operand (the product is A * B^T in KC/KC form)
  double* C; long ldc;                      // row-major M x N
  double alpha, beta;
  int flags;
  // Optional "staircase" structure (device arrays, one int per 128-wide tile, nullptr = off).
  // stair_n[t]: the columns of n-tile t (of the B operand / of C) are zero for global rows
  //   < stair_n[t]: tiles with stair_row0 + m0 + 128 <= stair_n[t] are skipped and k starts at
  //   stair_n[t] - stair_k0.   stair_m[t]: same for the A operand of m-tile t (k-range only).
  const int* stair_n;
  const int* stair_m;
  int stair_row0;   // global row index of C's row 0
  int stair_k0;     // global index of k = 0
};

struct Workspace;  // api.cu

// C = alpha * A(M x K) * B(N x K)^T + beta * C on `stream`.  All pointers device, 16-byte
// aligned, lda/ldb/ldc even.  Returns a MUNK_* status.
// `scratch` selects the split-K buffer: 0 for the caller's stream, 1 for the workspace's
// look-ahead stream (two launches that may run concurrently must not share one).
int gemm_launch(Workspace* ws, const GemmDesc& g, cudaStream_t stream, int scratch = 0);

// Flops actually executed as useful work by gemm_launch for this descriptor (tile-granular).
double gemm_flops(const GemmDesc& g);

}  // namespace munk

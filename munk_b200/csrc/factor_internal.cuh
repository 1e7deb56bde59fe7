// Building blocks of the factorisations shared by factor.cu (one GPU) and dist.cu (N GPUs).
#pragma once
#include "common.cuh"
#include "dgemm.cuh"
#include "workspace.cuh"

#include <cstdlib>
#include <vector>

namespace munk {
namespace detail {

constexpr int LEAF = MUNK_LEAF;

struct Rec {
  Workspace* ws;
  cudaStream_t stream;
  double* linv;      // leaf inverses, block g at linv + g*128*128 (g = gidx / 128)
  int scratch;       // split-K buffer of `stream` (workspace.cuh: 0 caller, 1 chain, 2 panel, 3+ pool)
  const int* stair = nullptr;   // forward left solve: first non-zero row of each 128-column RHS tile
  int stair_base = 0;           // global row index of gidx = 0 (stair values are global rows)
};

// largest multiple of 128 that is <= n/2 rounded up, strictly inside (0, n)
inline int split_point(int n) {
  int h = (int)round_up((n + 1) / 2, LEAF);
  if (h >= n) h -= LEAF;
  return h;
}

inline int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// In-place Cholesky (+ leaf inverses into r.linv) of the n x n block whose first column is gidx.
int potrf_rec(Rec& r, double* A, long ld, int n, int gidx);
// The same for a diagonal block of the block-column drivers (w <= 1024): 512-column pieces go to
// the one-launch kernel of diag.cu, anything else to the recursion.
int potrf_diag(Rec& r, double* A, long ld, int w, int gidx);
// B (m x s, in place) <- B L^-T, L the s x s lower block whose first column is gidx.
int trsm_rec(Rec& r, double* B, long ldb, int m, const double* L, long ldl, int s, int gidx);
// C (m x nc; top nc x nc part diagonal: lower tiles only) -= X Y^T, K = k
int syrk_update(Rec& r, double* C, long ldc, int m, int nc, const double* X, const double* Y, long ldx, int k);
// C (m x nc, every tile) -= X Y^T, K = k, operands with separate leading dimensions
int gemm_update(Rec& r, double* C, long ldc, int m, int nc, const double* X, const double* Y, long ldx, int k);
// B (n x nrhs, in place) <- L^-1 B or L^-T B
int trsm_left_rec(Rec& r, const double* L, long ldl, int n, double* B, long ldb, int nrhs, int gidx,
                  bool transposed);

}  // namespace detail
}  // namespace munk

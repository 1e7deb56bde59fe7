// Recursive blocked Cholesky / triangular inverse on one GPU.
//
// Replaces the LAPACK calls behind munk/munk/munk.py:83 (np.linalg.inv of I + lam*L) and
// munk/munk/munk.py:87-88 (scipy.linalg.eigh + diag GEMM): A = L L^T is SPD (eigenvalues >= 1),
// so D = A^-1 = W^T W with W = L^-1, and C = W^T is an RKHS factor (C C^T = D) without an
// eigensolver.
//
// Right-looking recursion, all matrices row-major, lower triangle, in place:
//   potrf(A):  potrf(A11);  A21 <- A21 L11^-T (trsm);  A22 <- A22 - A21 A21^T (DMMA syrk);
//              potrf(A22)
//   trsm(B,L): B1 <- trsm(B1,L11);  B2 <- B2 - B1 L21^T (DMMA gemm);  B2 <- trsm(B2,L22)
//   trtri(L):  W11 = trtri(L11); W22 = trtri(L22); W21 = -W22 (L21 W11)   (two DMMA trmm)
// Every split is a multiple of 128, so the recursion bottoms out on aligned 128x128 diagonal
// blocks.  One CTA factors such a block in shared memory and also inverts it; the inverse is
// kept (Workspace::linv) and turns the leaf trsm into a DMMA product and the leaf trtri
// into a copy.
#include "dgemm.cuh"
#include "factor.cuh"
#include "leaf.cuh"
#include "workspace.cuh"

#include <algorithm>

namespace munk {

namespace {

constexpr int LEAF = MUNK_LEAF;

// dst (nb x nb block inside a row-major matrix) <- src (128-ld leaf inverse)
__global__ void copy_leaf_kernel(double* __restrict__ dst, long ldd, const double* __restrict__ src,
                                 int nb) {
  for (int idx = threadIdx.x + blockIdx.x * blockDim.x; idx < nb * LEAF; idx += blockDim.x * gridDim.x) {
    const int r = idx / LEAF, c = idx % LEAF;
    if (c < nb) dst[(long)r * ldd + c] = src[idx];
  }
}

// largest multiple of 128 that is <= n/2 rounded up, strictly inside (0, n)
inline int split_point(int n) {
  int h = (int)round_up((n + 1) / 2, LEAF);
  if (h >= n) h -= LEAF;
  return h;
}

struct Rec {
  Workspace* ws;
  cudaStream_t stream;
  double* linv;      // leaf inverses of this matrix, block g at linv + g*128*128
};

int leaf_potrf(Rec& r, double* A, long ld, int nb, int gidx) {
  MUNK_TRY(launch_potrf_leaf(A, ld, nb, r.linv + (long)(gidx / LEAF) * LEAF * LEAF, gidx, r.ws->info,
                             r.stream));
  r.ws->launches++;
  r.ws->flops += (double)nb * nb * nb / 3.0;
  return MUNK_OK;
}

// B (m x s, in place) <- B * L^-T where L is the s x s lower block whose first column is gidx.
int trsm_rec(Rec& r, double* B, long ldb, int m, const double* L, long ldl, int s, int gidx) {
  if (m <= 0) return MUNK_OK;
  if (s <= LEAF) {
    GemmDesc g{};
    g.M = m; g.N = s; g.K = s;
    g.A = B; g.lda = ldb; g.a_layout = OP_KC;
    g.B = r.linv + (long)(gidx / LEAF) * LEAF * LEAF; g.ldb = LEAF; g.b_layout = OP_KC;
    g.C = B; g.ldc = ldb;     // in place: N <= 128, so each CTA owns its rows completely
    g.alpha = 1.0; g.beta = 0.0; g.flags = GF_KHI_N;
    return gemm_launch(r.ws, g, r.stream);
  }
  const int s1 = split_point(s), s2 = s - s1;
  MUNK_TRY(trsm_rec(r, B, ldb, m, L, ldl, s1, gidx));
  GemmDesc g{};
  g.M = m; g.N = s2; g.K = s1;
  g.A = B; g.lda = ldb; g.a_layout = OP_KC;
  g.B = L + (long)s1 * ldl; g.ldb = ldl; g.b_layout = OP_KC;     // L21
  g.C = B + s1; g.ldc = ldb;
  g.alpha = -1.0; g.beta = 1.0; g.flags = 0;
  MUNK_TRY(gemm_launch(r.ws, g, r.stream));
  return trsm_rec(r, B + s1, ldb, m, L + (long)s1 * ldl + s1, ldl, s2, gidx + s1);
}

int potrf_rec(Rec& r, double* A, long ld, int n, int gidx) {
  if (n <= LEAF) return leaf_potrf(r, A, ld, n, gidx);
  const int n1 = split_point(n), n2 = n - n1;
  MUNK_TRY(potrf_rec(r, A, ld, n1, gidx));
  double* A21 = A + (long)n1 * ld;
  double* A22 = A21 + n1;
  MUNK_TRY(trsm_rec(r, A21, ld, n2, A, ld, n1, gidx));
  GemmDesc g{};
  g.M = n2; g.N = n2; g.K = n1;
  g.A = A21; g.lda = ld; g.a_layout = OP_KC;
  g.B = A21; g.ldb = ld; g.b_layout = OP_KC;
  g.C = A22; g.ldc = ld;
  g.alpha = -1.0; g.beta = 1.0; g.flags = GF_TILE_LOWER;
  MUNK_TRY(gemm_launch(r.ws, g, r.stream));
  return potrf_rec(r, A22, ld, n2, gidx + n1);
}

int trtri_rec(Rec& r, double* L, long ld, int n, int gidx) {
  if (n <= LEAF) {
    copy_leaf_kernel<<<16, 256, 0, r.stream>>>(L, ld, r.linv + (long)(gidx / LEAF) * LEAF * LEAF, n);
    MUNK_CUDA_TRY(cudaGetLastError());
    r.ws->launches++;
    return MUNK_OK;
  }
  const int n1 = split_point(n), n2 = n - n1;
  MUNK_TRY(trtri_rec(r, L, ld, n1, gidx));
  MUNK_TRY(trtri_rec(r, L + (long)n1 * ld + n1, ld, n2, gidx + n1));
  double* L21 = L + (long)n1 * ld;
  const long ldt = round_up(n1, 16);
  // T = L21 * W11   (W11 lower: k >= column)
  GemmDesc g{};
  g.M = n2; g.N = n1; g.K = n1;
  g.A = L21; g.lda = ld; g.a_layout = OP_KC;
  g.B = L; g.ldb = ld; g.b_layout = OP_MC;
  g.C = r.ws->tmp; g.ldc = ldt;
  g.alpha = 1.0; g.beta = 0.0; g.flags = GF_KLO_N;
  MUNK_TRY(gemm_launch(r.ws, g, r.stream));
  // W21 = -W22 * T  (W22 lower: k <= row)
  GemmDesc h{};
  h.M = n2; h.N = n1; h.K = n2;
  h.A = L + (long)n1 * ld + n1; h.lda = ld; h.a_layout = OP_KC;
  h.B = r.ws->tmp; h.ldb = ldt; h.b_layout = OP_MC;
  h.C = L21; h.ldc = ld;
  h.alpha = -1.0; h.beta = 0.0; h.flags = GF_KHI_M;
  return gemm_launch(r.ws, h, r.stream);
}

}  // namespace

size_t linv_bytes_for(int n) { return (size_t)((n + LEAF - 1) / LEAF) * LEAF * LEAF * 8; }

int potrf_lower(Workspace* ws, int n, double* A, long ld, double* linv, cudaStream_t stream) {
  if (n <= 0) return MUNK_OK;
  Rec r{ws, stream, linv};
  return potrf_rec(r, A, ld, n, 0);
}

int trtri_lower(Workspace* ws, int n, double* L, long ld, double* linv, cudaStream_t stream) {
  if (n <= 0) return MUNK_OK;
  if (n > LEAF) {
    const int n1 = split_point(n);
    const size_t need = (size_t)(n - n1) * round_up(n1, 16) * 8;
    // the top-level temporary is the largest one of the recursion (n2 <= n1 at every level)
    MUNK_TRY(ws->ensure(&ws->tmp, &ws->tmp_bytes, std::max(need, (size_t)n1 * round_up(n1, 16) * 8)));
  }
  Rec r{ws, stream, linv};
  return trtri_rec(r, L, ld, n, 0);
}

}  // namespace munk

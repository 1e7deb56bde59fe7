// Blocked Cholesky / triangular inverse on one GPU.
//
// Replaces the LAPACK calls behind munk/munk/munk.py:83 (np.linalg.inv of I + lam*L) and
// munk/munk/munk.py:87-88 (scipy.linalg.eigh + diag GEMM): A = L L^T is SPD (eigenvalues >= 1),
// so D = A^-1 = W^T W with W = L^-1, and C = W^T is an RKHS factor (C C^T = D) without an
// eigensolver.
//
// All matrices row-major, lower triangle, in place.  Building blocks (recursive, every split a
// multiple of 128, so they bottom out on aligned 128x128 diagonal blocks):
//   potrf_rec(A):  potrf(A11);  A21 <- A21 L11^-T (trsm);  A22 <- A22 - A21 A21^T (DMMA syrk);
//                  potrf(A22)
//   trsm_rec(B,L): B1 <- trsm(B1,L11);  B2 <- B2 - B1 L21^T (DMMA gemm);  B2 <- trsm(B2,L22)
//   trtri_rec(L):  W11 = trtri(L11); W22 = trtri(L22); W21 = -W22 (L21 W11)   (two DMMA trmm)
// One CTA factors a 128x128 diagonal block and also inverts it (leaf.cu); the inverse is kept
// (`linv`) and turns the leaf trsm into a DMMA product and the leaf trtri into a copy.
//
// Large matrices use the right-looking blocked driver potrf_lookahead(): panels of 512 columns,
// the panel work (look-ahead update of the next block column, its recursive factorisation and
// panel solve: many small, latency-bound launches) runs on a high-priority side stream while the
// caller's stream applies the previous panel to the rest of the trailing matrix (one big DMMA
// syrk per panel), so the latency-bound part is hidden behind tensor-core work.
#include "dgemm.cuh"
#include "factor.cuh"
#include "leaf.cuh"
#include "workspace.cuh"

#include <algorithm>
#include <vector>

namespace munk {

namespace {

constexpr int LEAF = MUNK_LEAF;
constexpr int PANEL = 512;          // block-column width of the look-ahead driver ...
constexpr int PANEL_SMALL = 256;    // ... once fewer than PANEL_SWITCH columns remain
constexpr int PANEL_SWITCH = 8192;

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
  int scratch;       // split-K buffer of `stream` (0 caller's stream, 1 side stream)
  const int* stair = nullptr;   // forward left solve: first non-zero row of each 128-column RHS tile
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
    return gemm_launch(r.ws, g, r.stream, r.scratch);
  }
  const int s1 = split_point(s), s2 = s - s1;
  MUNK_TRY(trsm_rec(r, B, ldb, m, L, ldl, s1, gidx));
  GemmDesc g{};
  g.M = m; g.N = s2; g.K = s1;
  g.A = B; g.lda = ldb; g.a_layout = OP_KC;
  g.B = L + (long)s1 * ldl; g.ldb = ldl; g.b_layout = OP_KC;     // L21
  g.C = B + s1; g.ldc = ldb;
  g.alpha = -1.0; g.beta = 1.0; g.flags = 0;
  MUNK_TRY(gemm_launch(r.ws, g, r.stream, r.scratch));
  return trsm_rec(r, B + s1, ldb, m, L + (long)s1 * ldl + s1, ldl, s2, gidx + s1);
}

// C (lower tiles of an m x nc block whose top nc x nc part is diagonal) -= X Y^T, K = k
int syrk_update(Rec& r, double* C, long ldc, int m, int nc, const double* X, const double* Y, long ldx,
                int k) {
  GemmDesc g{};
  g.M = m; g.N = nc; g.K = k;
  g.A = X; g.lda = ldx; g.a_layout = OP_KC;
  g.B = Y; g.ldb = ldx; g.b_layout = OP_KC;
  g.C = C; g.ldc = ldc;
  g.alpha = -1.0; g.beta = 1.0; g.flags = GF_TILE_LOWER;
  return gemm_launch(r.ws, g, r.stream, r.scratch);
}

int potrf_rec(Rec& r, double* A, long ld, int n, int gidx) {
  if (n <= LEAF) return leaf_potrf(r, A, ld, n, gidx);
  const int n1 = split_point(n), n2 = n - n1;
  MUNK_TRY(potrf_rec(r, A, ld, n1, gidx));
  double* A21 = A + (long)n1 * ld;
  double* A22 = A21 + n1;
  MUNK_TRY(trsm_rec(r, A21, ld, n2, A, ld, n1, gidx));
  MUNK_TRY(syrk_update(r, A22, ld, n2, n2, A21, A21, ld, n1));
  return potrf_rec(r, A22, ld, n2, gidx + n1);
}

// Right-looking blocked Cholesky with one panel of look-ahead on the side stream.
//   step j (panel j = columns [c0, c0+w) already factored, event panel[j]):
//     side stream : wait upd[j-1];  block column j+1 -= panel j contributions (look-ahead);
//                   factor panel j+1 (potrf_rec of its diagonal block + trsm_rec below);
//                   record panel[j+1]
//     main stream : wait panel[j];  columns beyond block column j+1 -= panel j contributions
//                   (one lower-tile DMMA syrk, K = w);  record upd[j]
int potrf_lookahead(Workspace* ws, int n, double* A, long ld, double* linv, cudaStream_t main) {
  // Panel boundaries: 512 columns while the trailing matrix is large (each 128x128 tile of the
  // trailing syrk then runs 32 k-iterations and amortises its prologue/epilogue), 256 columns
  // once it is small, where the panel chain on the side stream is the critical path and a
  // shorter chain (and shorter main-stream tiles to wait behind) matters more.  Measured on
  // B200: n = 8192 13.3 -> 12.3 ms with 256 throughout, n = 20000 107 -> 127 ms.
  std::vector<int> start;
  for (int c = 0; c < n;) {
    start.push_back(c);
    c += (n - c > PANEL_SWITCH) ? PANEL : PANEL_SMALL;
  }
  start.push_back(n);
  const int nblk = (int)start.size() - 1;
  MUNK_TRY(ws->ensure_side(2 * (size_t)nblk + 2));
  MUNK_TRY(ws->ensure_splitk(0, SPLITK_RESERVE));
  MUNK_TRY(ws->ensure_splitk(1, SPLITK_RESERVE));
  cudaStream_t side = ws->side;
  cudaEvent_t* panel_ev = ws->events.data();          // [nblk]
  cudaEvent_t* upd_ev = ws->events.data() + nblk;     // [nblk]
  cudaEvent_t fork_ev = ws->events[2 * nblk];
  Rec rs{ws, side, linv, 1}, rm{ws, main, linv, 0};

  auto factor_panel = [&](int c0, int w) -> int {
    double* diag = A + (long)c0 * ld + c0;
    MUNK_TRY(potrf_rec(rs, diag, ld, w, c0));
    const int below = n - (c0 + w);
    if (below > 0) MUNK_TRY(trsm_rec(rs, diag + (long)w * ld, ld, below, diag, ld, w, c0));
    return MUNK_OK;
  };

  MUNK_CUDA_TRY(cudaEventRecord(fork_ev, main));
  MUNK_CUDA_TRY(cudaStreamWaitEvent(side, fork_ev, 0));
  MUNK_TRY(factor_panel(0, start[1]));
  MUNK_CUDA_TRY(cudaEventRecord(panel_ev[0], side));

  // MUNK_TRACE_POTRF=1: per-step device timeline (diagnostics only; synchronises at the end)
  const bool trace = getenv("MUNK_TRACE_POTRF") != nullptr;
  std::vector<cudaEvent_t> tm0, tm1, ts0, ts1, ts2;
  auto tick = [&](std::vector<cudaEvent_t>& v, cudaStream_t s) {
    if (!trace) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, s);
    v.push_back(e);
  };

  int last_panel = 0;
  for (int j = 0; j + 1 < nblk; ++j) {
    const int c0 = start[j], w = start[j + 1] - c0;
    const int r1 = c0 + w;                            // first row/column of block column j+1
    const int w1 = start[j + 2] - r1;
    const int r2 = r1 + w1;                           // first column the main stream updates
    const double* pan = A + (long)r1 * ld + c0;       // panel j rows >= r1
    // side: look-ahead update of block column j+1, then its factorisation
    if (j > 0) MUNK_CUDA_TRY(cudaStreamWaitEvent(side, upd_ev[j - 1], 0));
    tick(ts0, side);
    MUNK_TRY(syrk_update(rs, A + (long)r1 * ld + r1, ld, n - r1, w1, pan, pan, ld, w));
    tick(ts1, side);
    MUNK_TRY(factor_panel(r1, w1));
    tick(ts2, side);
    MUNK_CUDA_TRY(cudaEventRecord(panel_ev[j + 1], side));
    last_panel = j + 1;
    // main: the rest of the trailing matrix
    MUNK_CUDA_TRY(cudaStreamWaitEvent(main, panel_ev[j], 0));
    tick(tm0, main);
    if (r2 < n) {
      const double* pan2 = A + (long)r2 * ld + c0;    // panel j rows >= r2
      MUNK_TRY(syrk_update(rm, A + (long)r2 * ld + r2, ld, n - r2, n - r2, pan2, pan2, ld, w));
    }
    tick(tm1, main);
    MUNK_CUDA_TRY(cudaEventRecord(upd_ev[j], main));
  }
  MUNK_CUDA_TRY(cudaStreamWaitEvent(main, panel_ev[last_panel], 0));   // join
  if (trace) {
    cudaStreamSynchronize(main);
    fprintf(stderr, "potrf_lookahead n=%d: step rem | main: start dur | side: start lookahead panel (ms)\n", n);
    for (size_t j = 0; j < tm0.size(); ++j) {
      float a = 0, b = 0, c = 0, d = 0, e = 0;
      cudaEventElapsedTime(&a, tm0[0], tm0[j]);
      cudaEventElapsedTime(&b, tm0[j], tm1[j]);
      cudaEventElapsedTime(&c, tm0[0], ts0[j]);
      cudaEventElapsedTime(&d, ts0[j], ts1[j]);
      cudaEventElapsedTime(&e, ts1[j], ts2[j]);
      fprintf(stderr, "  %3zu %6d | %8.3f %7.3f | %8.3f %7.3f %7.3f\n", j, n - start[j + 1], a, b, c, d, e);
    }
    for (auto* v : {&tm0, &tm1, &ts0, &ts1, &ts2})
      for (cudaEvent_t ev : *v) cudaEventDestroy(ev);
  }
  return MUNK_OK;
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
  MUNK_TRY(gemm_launch(r.ws, g, r.stream, r.scratch));
  // W21 = -W22 * T  (W22 lower: k <= row)
  GemmDesc h{};
  h.M = n2; h.N = n1; h.K = n2;
  h.A = L + (long)n1 * ld + n1; h.lda = ld; h.a_layout = OP_KC;
  h.B = r.ws->tmp; h.ldb = ldt; h.b_layout = OP_MC;
  h.C = L21; h.ldc = ld;
  h.alpha = -1.0; h.beta = 0.0; h.flags = GF_KHI_M;
  return gemm_launch(r.ws, h, r.stream, r.scratch);
}

// Left solves with the Cholesky factor, in place on the row-major n x r block B:
//   forward : L X = B        X1 = fwd(L11,B1); B2 -= L21 X1;   X2 = fwd(L22,B2)
//   backward: L^T X = B      X2 = bwd(L22,B2); B1 -= L21^T X2; X1 = bwd(L11,B1)
// Leaves multiply by the stored inverse of the 128x128 diagonal block (M <= 128, so a CTA reads
// and writes only its own column strip of B: safe in place).
int trsm_left_rec(Rec& r, const double* L, long ldl, int n, double* B, long ldb, int nrhs, int gidx,
                  bool transposed) {
  if (n <= LEAF) {
    GemmDesc g{};
    g.M = n; g.N = nrhs; g.K = n;
    g.A = r.linv + (long)(gidx / LEAF) * LEAF * LEAF; g.lda = LEAF;
    g.a_layout = transposed ? OP_MC : OP_KC;
    g.B = B; g.ldb = ldb; g.b_layout = OP_MC;
    g.C = B; g.ldc = ldb;
    g.alpha = 1.0; g.beta = 0.0; g.flags = transposed ? GF_KLO_M : GF_KHI_M;
    if (!transposed && r.stair) { g.stair_n = r.stair; g.stair_row0 = gidx; g.stair_k0 = gidx; }
    return gemm_launch(r.ws, g, r.stream, r.scratch);
  }
  const int n1 = split_point(n), n2 = n - n1;
  const double* L21 = L + (long)n1 * ldl;
  const double* L22 = L21 + n1;
  double* B2 = B + (long)n1 * ldb;
  GemmDesc g{};
  g.alpha = -1.0; g.beta = 1.0; g.flags = 0;
  g.A = L21; g.lda = ldl;
  g.ldb = ldb; g.b_layout = OP_MC; g.ldc = ldb; g.N = nrhs;
  if (!transposed) {
    MUNK_TRY(trsm_left_rec(r, L, ldl, n1, B, ldb, nrhs, gidx, false));
    g.M = n2; g.K = n1; g.a_layout = OP_KC; g.B = B; g.C = B2;          // B2 -= L21 X1
    if (r.stair) { g.stair_n = r.stair; g.stair_row0 = gidx + n1; g.stair_k0 = gidx; }
    MUNK_TRY(gemm_launch(r.ws, g, r.stream, r.scratch));
    return trsm_left_rec(r, L22, ldl, n2, B2, ldb, nrhs, gidx + n1, false);
  }
  MUNK_TRY(trsm_left_rec(r, L22, ldl, n2, B2, ldb, nrhs, gidx + n1, true));
  g.M = n1; g.K = n2; g.a_layout = OP_MC; g.B = B2; g.C = B;            // B1 -= L21^T X2
  MUNK_TRY(gemm_launch(r.ws, g, r.stream, r.scratch));
  return trsm_left_rec(r, L, ldl, n1, B, ldb, nrhs, gidx, true);
}

}  // namespace

int trsm_left(Workspace* ws, int n, const double* L, long ld, const double* linv, int transposed,
              int nrhs, double* B, long ldb, const int* stair, cudaStream_t stream) {
  if (n <= 0 || nrhs <= 0) return MUNK_OK;
  Rec r{ws, stream, const_cast<double*>(linv), 0};
  r.stair = transposed ? nullptr : stair;
  return trsm_left_rec(r, L, ld, n, B, ldb, nrhs, 0, transposed != 0);
}

int trsm_right(Workspace* ws, int m, double* B, long ldb, const double* L, long ldl, int s,
               const double* linv, cudaStream_t stream) {
  if (m <= 0 || s <= 0) return MUNK_OK;
  Rec r{ws, stream, const_cast<double*>(linv), 0};
  return trsm_rec(r, B, ldb, m, L, ldl, s, 0);
}

size_t linv_bytes_for(int n) { return (size_t)((n + LEAF - 1) / LEAF) * LEAF * LEAF * 8; }

int potrf_lower(Workspace* ws, int n, double* A, long ld, double* linv, cudaStream_t stream) {
  if (n <= 0) return MUNK_OK;
  if (n > 2 * PANEL) return potrf_lookahead(ws, n, A, ld, linv, stream);
  Rec r{ws, stream, linv, 0};
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
  Rec r{ws, stream, linv, 0};
  return trtri_rec(r, L, ld, n, 0);
}

}  // namespace munk

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
// Large matrices use the right-looking blocked driver potrf_lookahead(): block columns of
// 1024/512/256 columns, scheduled critical path first on three streams (see its comment).
#include "dgemm.cuh"
#include "factor.cuh"
#include "factor_internal.cuh"
#include "leaf.cuh"
#include "workspace.cuh"

#include <algorithm>
#include <vector>

namespace munk {

namespace detail {

constexpr int PANEL = 512;          // default block-column width of the right-looking driver
constexpr int PANEL_SMALL = 256;    // ... near the end of the matrix (see panel_starts)

// dst (nb x nb block inside a row-major matrix) <- src (128-ld leaf inverse)
__global__ void copy_leaf_kernel(double* __restrict__ dst, long ldd, const double* __restrict__ src,
                                 int nb) {
  for (int idx = threadIdx.x + blockIdx.x * blockDim.x; idx < nb * LEAF; idx += blockDim.x * gridDim.x) {
    const int r = idx / LEAF, c = idx % LEAF;
    if (c < nb) dst[(long)r * ldd + c] = src[idx];
  }
}

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

int potrf_diag(Rec& r, double* A, long ld, int w, int gidx) {
  static const bool fused = env_int("MUNK_FUSED_DIAG", 1) != 0;
  if (!fused || w % LEAF != 0 || w < 2 * LEAF) return potrf_rec(r, A, ld, w, gidx);
  if (w <= 4 * LEAF) {
    MUNK_TRY(launch_potrf_diag(A, ld, w, r.linv + (long)(gidx / LEAF) * LEAF * LEAF, r.stair_base + gidx,
                               r.ws->info, reinterpret_cast<unsigned*>(r.ws->info + 4), r.stream));
    r.ws->launches += 2;
    r.ws->flops += (double)w * w * w / 3.0;
    return MUNK_OK;
  }
  // wider (the 1024-column panels of the early block columns): two fused halves
  const int n1 = 4 * LEAF, n2 = w - n1;
  MUNK_TRY(potrf_diag(r, A, ld, n1, gidx));
  double* A21 = A + (long)n1 * ld;
  double* A22 = A21 + n1;
  MUNK_TRY(launch_trsm_strip(A21, ld, n2, A, ld, n1, r.linv + (long)(gidx / LEAF) * LEAF * LEAF, r.stream));
  r.ws->launches++;
  r.ws->flops += (double)n2 * n1 * n1;
  MUNK_TRY(syrk_update(r, A22, ld, n2, n2, A21, A21, ld, n1));
  return potrf_diag(r, A22, ld, n2, gidx + n1);
}

// C (m x nc) -= X (m x k) Y (nc x k)^T, every tile (rows strictly below the diagonal block).
int gemm_update(Rec& r, double* C, long ldc, int m, int nc, const double* X, const double* Y, long ldx,
                int k) {
  if (m <= 0 || nc <= 0) return MUNK_OK;
  GemmDesc g{};
  g.M = m; g.N = nc; g.K = k;
  g.A = X; g.lda = ldx; g.a_layout = OP_KC;
  g.B = Y; g.ldb = ldx; g.b_layout = OP_KC;
  g.C = C; g.ldc = ldc;
  g.alpha = -1.0; g.beta = 1.0; g.flags = 0;
  return gemm_launch(r.ws, g, r.stream, r.scratch);
}

// Block-column boundaries of the right-looking driver: wide panels while the trailing matrix is
// large (each 128x128 tile of the trailing update then runs K/16 k-iterations and amortises its
// ~5.5 us of prologue/epilogue), narrower ones towards the end, where the diagonal-block chain is
// the critical path.  All multiples of 128.
std::vector<int> panel_starts(int n) {
  static const int big = env_int("MUNK_PANEL_BIG", 1024), big_until = env_int("MUNK_PANEL_BIG_UNTIL", 12288);
  static const int mid = env_int("MUNK_PANEL", PANEL), small = env_int("MUNK_PANEL_SMALL", PANEL_SMALL);
  static const int small_from = env_int("MUNK_PANEL_SWITCH", 4096);
  std::vector<int> start;
  for (int c = 0; c < n;) {
    start.push_back(c);
    const int rem = n - c;
    c += rem > big_until ? big : (rem > small_from ? mid : small);
  }
  start.push_back(n);
  return start;
}

// Right-looking blocked Cholesky, critical path first.  Block column c has diagonal rows
// D(c) = [s_c, s_c+1), next rows N(c) = [s_c+1, s_c+2) and rest R(c) = [s_c+2, n); P_c is its
// finished panel.  Tasks (all "-=" below are DMMA products with K = width of the earlier panel):
//   LA_D(c)  A[D(c),c] -= P_c-1[D(c)] P_c-1[D(c)]^T      LA_N/LA_R(c) the same for rows N(c)/R(c)
//   F(c)     Cholesky (+ leaf inverses) of the diagonal block
//   TS_N(c)  A[N(c),c] <- A[N(c),c] L_cc^-T               TS_R(c) the same for rows R(c)
//   T(c)     A[R(c), columns >= s_c+2] -= P_c[R(c)] P_c[R(c)]^T   (the bulk of the flops)
// The only serial chain of the factorisation is F(c) -> TS_N(c) -> LA_D(c+1) -> F(c+1): three
// small tasks per block column.  It gets its own highest-priority stream (`side`); everything
// else that feeds it (TS_R, LA_N, LA_R: tall GEMMs over the rows below) runs on a second stream
// one priority level lower (`side2`), rows needed first launched first, and overlaps the
// latency-bound F(c); the caller's stream runs T(c).  Round 1 had LA(whole column), F and
// TS(whole column) in series on one stream: ~1.3 ms per block column against ~0.4 ms now.
//   side : [upd c-2] LA_D(c)  F(c)  (diag c)  [lan c] TS_N(c)  (tsn c)
//   side2: [diag c] TS(N(c+1) rows)  [tsn c][upd c-1] LA_N(c+1) (lan c+1)  TS(rest)  (panel c)  LA_R(c+1)
//   main : [panel c] T(c) (upd c)                         [x] = wait event, (x) = record event
struct PotrfPlan {
  Workspace* ws;
  int n;
  double* A;
  long ld;
  double* linv;
  cudaStream_t main;
  std::vector<int> start;
  int nblk = 0;
  bool trace = false;
  std::vector<cudaEvent_t> tr;

  PotrfPlan(Workspace* ws_, int n_, double* A_, long ld_, double* linv_, cudaStream_t main_)
      : ws(ws_), n(n_), A(A_), ld(ld_), linv(linv_), main(main_), start(panel_starts(n_)) {
    nblk = (int)start.size() - 1;
  }
  int s(int c) const { return start[std::min(c, nblk)]; }
  double* at(int row, int col) const { return A + (long)row * ld + col; }
  // flops of the trailing update T(c): the weight used to interleave two factorisations
  double work(int c) const {
    const double m = n - s(c + 2);
    return m * m * (s(c + 1) - s(c));
  }
  cudaEvent_t* ev(int which) const { return ws->events.data() + (size_t)which * nblk; }

  int begin() {
    MUNK_TRY(ws->ensure_side(5 * (size_t)nblk + 2));
    for (int i = 0; i < 3; ++i) MUNK_TRY(ws->ensure_splitk(i, SPLITK_RESERVE));
    cudaEvent_t fork_ev = ws->events[5 * nblk];
    MUNK_CUDA_TRY(cudaEventRecord(fork_ev, main));
    MUNK_CUDA_TRY(cudaStreamWaitEvent(ws->side, fork_ev, 0));
    MUNK_CUDA_TRY(cudaStreamWaitEvent(ws->side2, fork_ev, 0));
    return MUNK_OK;
  }

  // A[r0:r1, block column c] -= P_c-1[r0:r1] P_c-1[D(c)]^T
  int la(Rec& r, int r0, int r1, int c) {
    const int k = s(c) - s(c - 1), w = s(c + 1) - s(c);
    if (r0 == s(c)) return syrk_update(r, at(r0, s(c)), ld, r1 - r0, w, at(r0, s(c - 1)), at(s(c), s(c - 1)), ld, k);
    return gemm_update(r, at(r0, s(c)), ld, r1 - r0, w, at(r0, s(c - 1)), at(s(c), s(c - 1)), ld, k);
  }
  // A[r0:r1, block column c] <- . L_cc^-T.  Few rows (the chain's TS_N): one strip-solve launch
  // instead of the recursion's ~14 dependent small GEMMs.
  int ts(Rec& r, int r0, int r1, int c) {
    const int w = s(c + 1) - s(c);
    static const int strip_rows = env_int("MUNK_STRIP_ROWS", 1024);
    if (r1 - r0 > 0 && r1 - r0 <= strip_rows && w <= 512) {
      MUNK_TRY(launch_trsm_strip(at(r0, s(c)), ld, r1 - r0, at(s(c), s(c)), ld, w,
                                 linv + (long)(s(c) / LEAF) * LEAF * LEAF, r.stream));
      ws->launches++;
      ws->flops += (double)(r1 - r0) * w * w;
      return MUNK_OK;
    }
    return trsm_rec(r, at(r0, s(c)), ld, r1 - r0, at(s(c), s(c)), ld, w, s(c));
  }
  void tick(cudaStream_t st) {
    if (!trace) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    tr.push_back(e);
  }

  int column(int c) {
    cudaStream_t side = ws->side, side2 = ws->side2;
    cudaEvent_t *diag_ev = ev(0), *tsn_ev = ev(1), *lan_ev = ev(2), *panel_ev = ev(3), *upd_ev = ev(4);
    Rec rs{ws, side, linv, 1}, rp{ws, side2, linv, 2}, rm{ws, main, linv, 0};
    // ---- chain: LA_D(c), F(c), TS_N(c)
    if (c >= 2) MUNK_CUDA_TRY(cudaStreamWaitEvent(side, upd_ev[c - 2], 0));
    tick(side);
    if (c >= 1) MUNK_TRY(la(rs, s(c), s(c + 1), c));
    tick(side);
    MUNK_TRY(potrf_diag(rs, at(s(c), s(c)), ld, s(c + 1) - s(c), s(c)));
    MUNK_CUDA_TRY(cudaEventRecord(diag_ev[c], side));
    tick(side);
    if (c >= 1) MUNK_CUDA_TRY(cudaStreamWaitEvent(side, lan_ev[c], 0));
    if (c + 1 < nblk) MUNK_TRY(ts(rs, s(c + 1), s(c + 2), c));
    MUNK_CUDA_TRY(cudaEventRecord(tsn_ev[c], side));
    tick(side);
    // ---- panel stream: the rest of P_c, then the look-ahead updates of block column c + 1
    MUNK_CUDA_TRY(cudaStreamWaitEvent(side2, diag_ev[c], 0));
    if (c + 2 < nblk) MUNK_TRY(ts(rp, s(c + 2), s(c + 3), c));
    MUNK_CUDA_TRY(cudaStreamWaitEvent(side2, tsn_ev[c], 0));
    if (c >= 1) MUNK_CUDA_TRY(cudaStreamWaitEvent(side2, upd_ev[c - 1], 0));
    if (c + 2 < nblk) MUNK_TRY(la(rp, s(c + 2), s(c + 3), c + 1));
    if (c + 1 < nblk) MUNK_CUDA_TRY(cudaEventRecord(lan_ev[c + 1], side2));
    if (c + 3 < nblk) MUNK_TRY(ts(rp, s(c + 3), n, c));
    MUNK_CUDA_TRY(cudaEventRecord(panel_ev[c], side2));
    if (c + 3 < nblk) MUNK_TRY(la(rp, s(c + 3), n, c + 1));
    // ---- caller's stream: the trailing update beyond block column c + 1
    MUNK_CUDA_TRY(cudaStreamWaitEvent(main, panel_ev[c], 0));
    if (c + 2 < nblk) {
      const int r2 = s(c + 2);
      MUNK_TRY(syrk_update(rm, at(r2, r2), ld, n - r2, n - r2, at(r2, s(c)), at(r2, s(c)), ld, s(c + 1) - s(c)));
    }
    MUNK_CUDA_TRY(cudaEventRecord(upd_ev[c], main));
    return MUNK_OK;
  }

  int end() {
    cudaEvent_t join_ev = ws->events[5 * nblk + 1];
    MUNK_CUDA_TRY(cudaEventRecord(join_ev, ws->side));
    MUNK_CUDA_TRY(cudaStreamWaitEvent(main, join_ev, 0));           // side2 joined through panel_ev
    return MUNK_OK;
  }
};

int potrf_lookahead(Workspace* ws, int n, double* A, long ld, double* linv, cudaStream_t main) {
  PotrfPlan p(ws, n, A, ld, linv, main);
  // MUNK_TRACE_POTRF=1: per-column device timeline of the chain stream (diagnostics only)
  p.trace = getenv("MUNK_TRACE_POTRF") != nullptr;
  MUNK_TRY(p.begin());
  for (int c = 0; c < p.nblk; ++c) MUNK_TRY(p.column(c));
  MUNK_TRY(p.end());
  if (p.trace) {
    cudaStreamSynchronize(main);
    fprintf(stderr, "potrf n=%d: col rem | chain start, LA_D, F, wait+TS_N (ms)\n", n);
    for (int c = 0; c < p.nblk; ++c) {
      float t0 = 0, a = 0, b = 0, d = 0;
      cudaEventElapsedTime(&t0, p.tr[0], p.tr[4 * c]);
      cudaEventElapsedTime(&a, p.tr[4 * c], p.tr[4 * c + 1]);
      cudaEventElapsedTime(&b, p.tr[4 * c + 1], p.tr[4 * c + 2]);
      cudaEventElapsedTime(&d, p.tr[4 * c + 2], p.tr[4 * c + 3]);
      fprintf(stderr, "  %3d %6d | %8.3f %7.3f %7.3f %7.3f\n", c, n - p.start[c], t0, a, b, d);
    }
    for (cudaEvent_t e : p.tr) cudaEventDestroy(e);
  }
  return MUNK_OK;
}

// Two independent factorisations whose trailing updates share ONE stream, block columns
// interleaved in proportion to their work.  Each matrix keeps its own chain and panel streams
// (its own workspace).  Left to the hardware scheduler, two factorisations on two streams
// barely overlap (measured: together they take the sum of their separate times, 95 + 52 ms at
// 20 000 / 16 000 columns) because a ready grid of the earlier stream always goes first; in one
// stream the order is ours, and while one matrix waits for its latency-bound chain
// (LA_D -> F -> TS_N) the other one's trailing update keeps the SMs busy.  `lead` > 1 would let
// the second matrix run ahead of its share; measured at 20 000 / 16 000 columns: 137.7 ms with
// lead 1.0, 142.3 with 1.3, 144.9 with 2.0 (separately: 95.6 + 51.8).
int potrf_pair_lookahead(Workspace* wa, int na, double* A, long lda, double* linva, Workspace* wb, int nb,
                         double* B, long ldb, double* linvb, cudaStream_t main) {
  PotrfPlan pa(wa, na, A, lda, linva, main), pb(wb, nb, B, ldb, linvb, main);
  MUNK_TRY(pa.begin());
  MUNK_TRY(pb.begin());
  double ta = 0, tb = 0, da = 0, db = 0;
  for (int c = 0; c < pa.nblk; ++c) ta += pa.work(c) + 1.0;
  for (int c = 0; c < pb.nblk; ++c) tb += pb.work(c) + 1.0;
  static const double lead = env_int("MUNK_PAIR_LEAD_PCT", 100) / 100.0;
  int ca = 0, cb = 0;
  while (ca < pa.nblk || cb < pb.nblk) {
    const bool take_b = cb < pb.nblk && (ca >= pa.nblk || db / tb <= lead * da / ta);
    if (take_b) {
      MUNK_TRY(pb.column(cb));
      db += pb.work(cb) + 1.0;
      ++cb;
    } else {
      MUNK_TRY(pa.column(ca));
      da += pa.work(ca) + 1.0;
      ++ca;
    }
  }
  MUNK_TRY(pa.end());
  MUNK_TRY(pb.end());
  return MUNK_OK;
}

// Triangular inverse by recursive halving.  The two half-size sub-inversions are independent:
// down to `fork_depth` levels the second one runs on a pool stream (own split-K scratch, its own
// half of the temporary), so up to 8 subtrees are in flight and their small, latency-bound
// products and leaf copies overlap instead of queueing on one stream.
struct TrtriCtx {
  int next_pool = 0;        // pool streams handed out so far
  size_t next_event;        // events handed out so far (index into ws->events)
};

int trtri_rec(Rec& r, TrtriCtx& cx, double* L, long ld, int n, int gidx, double* tmp, size_t tmp_elems,
              int depth) {
  if (n <= LEAF) {
    copy_leaf_kernel<<<16, 256, 0, r.stream>>>(L, ld, r.linv + (long)(gidx / LEAF) * LEAF * LEAF, n);
    MUNK_CUDA_TRY(cudaGetLastError());
    r.ws->launches++;
    return MUNK_OK;
  }
  const int n1 = split_point(n), n2 = n - n1;
  static const int fork_depth = env_int("MUNK_TRTRI_FORK_DEPTH", 3);
  const bool fork = depth < fork_depth && n >= 1024 && cx.next_pool < POOL_STREAMS;
  if (fork) {
    const int pi = cx.next_pool++;
    Rec r2 = r;
    r2.stream = r.ws->pool[pi];
    r2.scratch = 3 + pi;
    MUNK_TRY(r.ws->ensure_splitk(r2.scratch, SPLITK_RESERVE));
    MUNK_TRY(r.ws->ensure_events(cx.next_event + 2));
    cudaEvent_t f = r.ws->events[cx.next_event++], j = r.ws->events[cx.next_event++];
    MUNK_CUDA_TRY(cudaEventRecord(f, r.stream));
    MUNK_CUDA_TRY(cudaStreamWaitEvent(r2.stream, f, 0));
    const size_t half = tmp_elems / 2 / 2 * 2;          // keeps 16-byte alignment
    MUNK_TRY(trtri_rec(r, cx, L, ld, n1, gidx, tmp, half, depth + 1));
    MUNK_TRY(trtri_rec(r2, cx, L + (long)n1 * ld + n1, ld, n2, gidx + n1, tmp + half, half, depth + 1));
    MUNK_CUDA_TRY(cudaEventRecord(j, r2.stream));
    MUNK_CUDA_TRY(cudaStreamWaitEvent(r.stream, j, 0));
  } else {
    MUNK_TRY(trtri_rec(r, cx, L, ld, n1, gidx, tmp, tmp_elems, depth + 1));
    MUNK_TRY(trtri_rec(r, cx, L + (long)n1 * ld + n1, ld, n2, gidx + n1, tmp, tmp_elems, depth + 1));
  }
  double* L21 = L + (long)n1 * ld;
  const long ldt = round_up(n1, 16);
  if ((size_t)n2 * ldt > tmp_elems) {
    set_last_error("trtri temporary too small (%d x %ld needs %zu elements, have %zu)", n2, ldt,
                   (size_t)n2 * ldt, tmp_elems);
    return -20;
  }
  // T = L21 * W11   (W11 lower: k >= column)
  GemmDesc g{};
  g.M = n2; g.N = n1; g.K = n1;
  g.A = L21; g.lda = ld; g.a_layout = OP_KC;
  g.B = L; g.ldb = ld; g.b_layout = OP_MC;
  g.C = tmp; g.ldc = ldt;
  g.alpha = 1.0; g.beta = 0.0; g.flags = GF_KLO_N;
  MUNK_TRY(gemm_launch(r.ws, g, r.stream, r.scratch));
  // W21 = -W22 * T  (W22 lower: k <= row)
  GemmDesc h{};
  h.M = n2; h.N = n1; h.K = n2;
  h.A = L + (long)n1 * ld + n1; h.lda = ld; h.a_layout = OP_KC;
  h.B = tmp; h.ldb = ldt; h.b_layout = OP_MC;
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
    if (!transposed && r.stair) { g.stair_n = r.stair; g.stair_row0 = r.stair_base + gidx; g.stair_k0 = r.stair_base + gidx; }
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
    if (r.stair) { g.stair_n = r.stair; g.stair_row0 = r.stair_base + gidx + n1; g.stair_k0 = r.stair_base + gidx; }
    MUNK_TRY(gemm_launch(r.ws, g, r.stream, r.scratch));
    return trsm_left_rec(r, L22, ldl, n2, B2, ldb, nrhs, gidx + n1, false);
  }
  MUNK_TRY(trsm_left_rec(r, L22, ldl, n2, B2, ldb, nrhs, gidx + n1, true));
  g.M = n1; g.K = n2; g.a_layout = OP_MC; g.B = B2; g.C = B;            // B1 -= L21^T X2
  MUNK_TRY(gemm_launch(r.ws, g, r.stream, r.scratch));
  return trsm_left_rec(r, L, ldl, n1, B, ldb, nrhs, gidx, true);
}

}  // namespace detail

using namespace detail;

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
  static const int strip_rows = env_int("MUNK_STRIP_ROWS", 1024);
  if (m <= strip_rows && s <= 512) {
    MUNK_TRY(launch_trsm_strip(B, ldb, m, L, ldl, s, linv, stream));
    ws->launches++;
    ws->flops += (double)m * s * s;
    return MUNK_OK;
  }
  Rec r{ws, stream, const_cast<double*>(linv), 0};
  return trsm_rec(r, B, ldb, m, L, ldl, s, 0);
}

size_t linv_bytes_for(int n) { return (size_t)((n + LEAF - 1) / LEAF) * LEAF * LEAF * 8; }

namespace {

int potrf_direct(Workspace* ws, int n, double* A, long ld, double* linv, cudaStream_t stream) {
  if (n > 2 * PANEL) return potrf_lookahead(ws, n, A, ld, linv, stream);
  Rec r{ws, stream, linv, 0};
  return potrf_rec(r, A, ld, n, 0);
}

int trtri_direct(Workspace* ws, int n, double* L, long ld, double* linv, cudaStream_t stream) {
  size_t elems = 0;
  if (n > LEAF) {
    // the top-level temporary (n2 x n1) is the largest of the recursion; a forked level gives
    // each half of it to one sub-inversion, whose own temporary is a quarter of that size
    const int n1 = split_point(n);
    elems = std::max((size_t)(n - n1) * round_up(n1, 16), (size_t)n1 * round_up(n1, 16));
    MUNK_TRY(ws->ensure(&ws->tmp, &ws->tmp_bytes, elems * 8));
    MUNK_TRY(ws->ensure_pool(stream));
    MUNK_TRY(ws->ensure_splitk(0, SPLITK_RESERVE));
  }
  Rec r{ws, stream, linv, 0};
  TrtriCtx cx;
  cx.next_event = 0;
  return trtri_rec(r, cx, L, ld, n, 0, ws->tmp, elems, 0);
}

constexpr int GRAPH_MIN_N = 4096;      // below this the direct path costs the host < 3 ms
constexpr size_t GRAPH_CACHE = 8;

// Runs `direct` (op 0 = potrf, 1 = trtri) either directly or as a cached CUDA graph.  First use
// of a (op, n, pointers) key: direct, which also performs every allocation and kernel-attribute
// call (none may happen inside a capture).  Second use: captured from `stream` - the side and
// pool streams join the capture through the fork event - instantiated and launched.  Later
// uses: one cudaGraphLaunch.  The temporary of trtri is part of the key's validity: it only
// grows, and growing it (a larger n) drops the cached graphs.
struct SecondKey { int n = 0; const void* a = nullptr; long ld = 0; const void* linv = nullptr; };

template <typename F>
int graph_or_direct(Workspace* ws, int op, int n, double* A, long ld, double* linv, cudaStream_t stream,
                    F direct, SecondKey k2 = SecondKey()) {
  static const bool off = getenv("MUNK_NO_GRAPH") != nullptr;
  if (off || n < GRAPH_MIN_N || getenv("MUNK_TRACE_POTRF")) return direct();
  // the legacy / per-thread default streams cannot be captured
  if (stream == nullptr || stream == cudaStreamLegacy || stream == cudaStreamPerThread) return direct();
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
    (void)cudaGetLastError();
    return direct();                       // the caller is capturing already: just record into it
  }
  Workspace::GraphEntry* hit = nullptr;
  for (auto& g : ws->graphs)
    if (g.op == op && g.n == n && g.a == A && g.ld == ld && g.linv == linv && g.n2 == k2.n && g.a2 == k2.a &&
        g.ld2 == k2.ld && g.linv2 == k2.linv)
      hit = &g;
  if (!hit) {
    if (ws->graphs.size() >= GRAPH_CACHE) {
      size_t old = 0;
      for (size_t i = 1; i < ws->graphs.size(); ++i)
        if (ws->graphs[i].stamp < ws->graphs[old].stamp) old = i;
      if (ws->graphs[old].exec) cudaGraphExecDestroy(ws->graphs[old].exec);
      ws->graphs.erase(ws->graphs.begin() + old);
    }
    ws->graphs.push_back({op, n, A, ld, linv, k2.n, k2.a, k2.ld, k2.linv, nullptr, 0, 0.0, ++ws->graph_clock});
    return direct();
  }
  hit->stamp = ++ws->graph_clock;
  if (!hit->exec) {
    const double* tmp_before = ws->tmp;
    const long l0 = ws->launches;
    const double f0 = ws->flops;
    MUNK_CUDA_TRY(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
    const int st = direct();
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(stream, &graph);
    if (st != MUNK_OK || e != cudaSuccess || !graph || ws->tmp != tmp_before) {
      if (getenv("MUNK_GRAPH_DEBUG"))
        fprintf(stderr, "munk: graph capture of op %d n=%d failed (status %d, %s)\n", op, n, st,
                cudaGetErrorString(e));
      if (graph) cudaGraphDestroy(graph);
      (void)cudaGetLastError();
      ws->launches = l0; ws->flops = f0;
      hit->op = -1;                        // never try this key again
      if (st != MUNK_OK) return st;
      return direct();
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) {
      if (getenv("MUNK_GRAPH_DEBUG"))
        fprintf(stderr, "munk: graph instantiation of op %d n=%d failed (%s)\n", op, n, cudaGetErrorString(ie));
      (void)cudaGetLastError();
      ws->launches = l0; ws->flops = f0;
      hit->op = -1;
      return direct();
    }
    if (getenv("MUNK_GRAPH_DEBUG"))
      fprintf(stderr, "munk: captured op %d n=%d into a graph (%ld launches)\n", op, n, ws->launches - l0);
    hit->exec = exec;
    hit->launches = ws->launches - l0;
    hit->flops = ws->flops - f0;
    ws->launches = l0; ws->flops = f0;
  }
  MUNK_CUDA_TRY(cudaGraphLaunch(hit->exec, stream));
  ws->launches += hit->launches;
  ws->flops += hit->flops;
  return MUNK_OK;
}

}  // namespace

int potrf_lower(Workspace* ws, int n, double* A, long ld, double* linv, cudaStream_t stream) {
  if (n <= 0) return MUNK_OK;
  return graph_or_direct(ws, 0, n, A, ld, linv, stream,
                         [&] { return potrf_direct(ws, n, A, ld, linv, stream); });
}

int trtri_lower(Workspace* ws, int n, double* L, long ld, double* linv, cudaStream_t stream) {
  if (n <= 0) return MUNK_OK;
  return graph_or_direct(ws, 1, n, L, ld, linv, stream,
                         [&] { return trtri_direct(ws, n, L, ld, linv, stream); });
}

int potrf_pair_lower(Workspace* wa, int na, double* A, long lda, double* linva, Workspace* wb, int nb,
                     double* B, long ldb, double* linvb, cudaStream_t stream) {
  if (na <= 2 * PANEL || nb <= 2 * PANEL || wa == wb) {          // nothing to interleave
    MUNK_TRY(potrf_lower(wa, na, A, lda, linva, stream));
    return potrf_lower(wb, nb, B, ldb, linvb, stream);
  }
  // launches / flops of both matrices are accounted to the first workspace (the graph's owner)
  return graph_or_direct(wa, 2, na, A, lda, linva, stream,
                         [&] {
                           const long lb = wb->launches;
                           const double fb = wb->flops;
                           const int st = potrf_pair_lookahead(wa, na, A, lda, linva, wb, nb, B, ldb, linvb, stream);
                           wa->launches += wb->launches - lb; wb->launches = lb;
                           wa->flops += wb->flops - fb; wb->flops = fb;
                           return st;
                         },
                         SecondKey{nb, B, ldb, linvb});
}

}  // namespace munk

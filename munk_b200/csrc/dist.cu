// Multi-GPU Cholesky with the forward substitution fused into the panel loop.
//
// One process per GPU; 1 x P block-column-cyclic ownership of the block columns (width 512,
// serpentine rounds, the same map as munk_b200/dist.py ColumnCyclic).  The factor lives in
// PANEL-MAJOR storage: block column c is one contiguous buffer
//     [ leaf inverses of its diagonal block | rows s_c .. n of the column, leading dimension w_c ]
// which is at once the owner's working storage (trailing updates accumulate into it, the
// factorisation runs in place) and every other rank's receive buffer: ncclBroadcast goes straight
// from and into it, no pack / unpack copies, no strided traffic.  A rank assembles and updates only
// the columns it owns (work and HBM traffic of A are 1/P per rank); after the loop every rank
// holds the whole factor, which is what makes the later stages exchange-free.
//
// Per block column c (s_c = first column, D(c) its rows, N(c) the next block's rows, R(c) the rest)
// the tasks and streams are those of the one-GPU driver (factor.cu potrf_lookahead), with the
// owner doing the column work and two broadcasts inserted:
//   chain  (highest priority)  owner: LA_D(c)  F(c)  TS_N(c)        all: bcast A(c) = [linv | D(c) N(c)]
//   panel  (next priority)     owner: TS_R(c)                        all: bcast B(c) = rows R(c)
//                              owner of c+1: LA_N(c+1), LA_R(c+1)
//   caller's stream            every rank: T(c) on its own block columns >= c+2
//   solve  (fused)             every rank: X[D(c)] <- L_cc^-1 X[D(c)];  X[below] -= P_c[below] X[D(c)]
// The serial chain per block column is LA_D -> F -> TS_N -> bcast A (a few MB); B(c) travels while
// the next owner factors its diagonal block.  The forward substitution L X = E of the columns of
// L^-1 this rank needs (its own block columns and the landmark columns, staircase right-hand side)
// used to be a separate phase after the factorisation (23 ms of 112 at 8 GPUs); here each panel is
// applied to X the moment it arrives, filling the SMs the latency-bound chain leaves idle.
//
// Reference call sites replaced: munk/munk/munk.py:83 (np.linalg.inv) and :87-88 (eigh factor) for
// a matrix sharded over GPUs.
#include "../../include/munk_b200.h"

#include "common.cuh"
#include "dgemm.cuh"
#include "factor_internal.cuh"
#include "kernels.cuh"
#include "leaf.cuh"
#include "workspace.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <new>
#include <vector>

using namespace munk;
using namespace munk::detail;

// ------------------------------------------------------------------------------------------------
// NCCL, bound at run time: the library is the one the process already loaded (torch ships
// libnccl.so.2), so a single-GPU user never needs it and nothing is linked against a second copy.
// ------------------------------------------------------------------------------------------------
namespace {

struct NcclApi {
  bool ok = false;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig_t*) = nullptr;   // optional
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi& nccl() {
  static NcclApi api = [] {
    NcclApi a;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);        // already in the process?
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return a;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
    a.CommInitRankConfig = (decltype(a.CommInitRankConfig))dlsym(h, "ncclCommInitRankConfig");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
    a.Broadcast = (decltype(a.Broadcast))dlsym(h, "ncclBroadcast");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.Broadcast && a.GetErrorString;
    return a;
  }();
  return api;
}

#define MUNK_ERR_NCCL 4000
#define MUNK_NCCL_TRY(expr)                                                                      \
  do {                                                                                           \
    ncclResult_t _r = (expr);                                                                    \
    if (_r != ncclSuccess) {                                                                     \
      munk::set_last_error("%s failed: %s (%s:%d)", #expr, nccl().GetErrorString(_r), __FILE__,  \
                           __LINE__);                                                            \
      return MUNK_ERR_NCCL + (int)_r;                                                            \
    }                                                                                            \
  } while (0)

constexpr int NB = 512;     // block-column width of the distributed layout

inline int owner_of(int j, int world) {
  const int r = j % world, rnd = j / world;
  return (rnd % 2 == 0) ? r : world - 1 - r;
}

// rows [0, rows) of one owned panel: zeros, and 1 + lam * L_ii on the diagonal of the first w rows
__global__ void __launch_bounds__(256)
fill_panel_kernel(int rows, int w, long ldp, int s_c, double lam, const int* __restrict__ deg,
                  const double* __restrict__ ldiag, double* __restrict__ P) {
  const int r = blockIdx.x;
  const int i = s_c + r;
  const double diag = __dadd_rn(1.0, __dmul_rn(lam, ldiag ? ldiag[i] : (double)deg[i]));
  double2* __restrict__ p = reinterpret_cast<double2*>(P + (long)r * ldp);
  const int npair = (int)(ldp >> 1);
  for (int c = blockIdx.y * blockDim.x + threadIdx.x; c < npair; c += gridDim.y * blockDim.x) {
    double2 v = make_double2(0.0, 0.0);
    if (r < w && c == (r >> 1)) {
      if (r & 1) v.y = diag; else v.x = diag;
    }
    p[c] = v;
  }
}

// lower-triangle entry of every edge into the panel of its column, if this rank owns it
__global__ void scatter_panel_kernel(int n, int nnz, const int* __restrict__ u, const int* __restrict__ v,
                                     const double* __restrict__ lval, double lam,
                                     const int* __restrict__ colblk, const int* __restrict__ start,
                                     const long* __restrict__ rows_off, const int* __restrict__ ldp,
                                     const unsigned char* __restrict__ owned, double* __restrict__ storage,
                                     int* __restrict__ bad) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  const int a = u[e], b = v[e];
  if ((unsigned)a >= (unsigned)n || (unsigned)b >= (unsigned)n) {
    atomicCAS(bad, 0, e + 1);
    return;
  }
  if (a == b) return;
  const int lo = min(a, b), hi = max(a, b);
  const int c = colblk[lo];
  if (!owned[c]) return;
  storage[rows_off[c] + (long)(hi - start[c]) * ldp[c] + (lo - start[c])] = __dmul_rn(lam, lval ? lval[e] : -1.0);
}

__global__ void degree_count_kernel(int n, int nnz, const int* __restrict__ u, const int* __restrict__ v,
                                    int* __restrict__ deg, int* __restrict__ bad) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < nnz) {
    const int a = u[e], b = v[e];
    if ((unsigned)a >= (unsigned)n || (unsigned)b >= (unsigned)n) {
      atomicCAS(bad, 0, e + 1);
      return;
    }
    if (a == b) return;
    atomicAdd(deg + a, 1);
    atomicAdd(deg + b, 1);
  }
}

// L[s_c + r][s_c + j] = panel[r][j] for j <= r-within-diagonal-block or any j below it
__global__ void unpack_panel_kernel(int rows, int w, long ldp, const double* __restrict__ P,
                                    double* __restrict__ L, long ld) {
  const int r = blockIdx.x;
  for (int j = threadIdx.x; j < w; j += blockDim.x)
    L[(long)r * ld + j] = (r >= w || j <= r) ? P[(long)r * ldp + j] : 0.0;
}

}  // namespace

struct munk_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
};

struct munk_dist {
  Workspace* ws = nullptr;
  munk_comm* comm = nullptr;        // chain communicator: the small piece A(c) of every panel
  munk_comm* bulk = nullptr;        // bulk communicator: the rest of the panel (never queues ahead of a piece A)
  int rank = 0, world = 1, n = 0, nblk = 0;
  std::vector<int> start;                 // nblk + 1
  std::vector<int> ldp;                   // leading dimension of panel c
  std::vector<long> linv_off, rows_off;   // element offsets into storage
  double* storage = nullptr;
  // device copies of the layout (assembly kernels)
  int* d_colblk = nullptr; int* d_start = nullptr; int* d_ldp = nullptr; long* d_rows_off = nullptr;
  unsigned char* d_owned = nullptr;
  // fused forward substitution
  double* X = nullptr; long ldx = 0; int ncols = 0; const int* stair = nullptr;
  cudaStream_t main = nullptr, solve = nullptr;
  cudaEvent_t solve_done = nullptr;
  int next = 0;                           // next block column of the running factorisation
  bool running = false;
  // MUNK_TRACE_DIST=1: device timeline of the chain / panel streams (diagnostics only)
  bool trace = false;
  std::vector<cudaEvent_t> tr;
  void tick(cudaStream_t st) {
    if (!trace) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    tr.push_back(e);
  }

  int w(int c) const { return start[c + 1] - start[c]; }
  int s(int c) const { return start[std::min(c, nblk)]; }
  bool own(int c) const { return c < nblk && owner_of(c, world) == rank; }
  double* linv(int c) const { return storage + linv_off[c]; }
  double* rows(int c, int global_row) const { return storage + rows_off[c] + (long)(global_row - start[c]) * ldp[c]; }
  long leaf_elems(int c) const { return (long)((w(c) + LEAF - 1) / LEAF) * LEAF * LEAF; }
};

namespace {

void layout(int n, std::vector<int>& start, std::vector<int>& ldp, std::vector<long>& linv_off,
            std::vector<long>& rows_off, long* total) {
  start.clear(); ldp.clear(); linv_off.clear(); rows_off.clear();
  for (int c = 0; c < n; c += NB) start.push_back(c);
  start.push_back(n);
  long off = 0;
  for (size_t c = 0; c + 1 < start.size(); ++c) {
    const int w = start[c + 1] - start[c];
    const long ld = round_up(w, 16);
    ldp.push_back((int)ld);
    linv_off.push_back(off);
    off += (long)((w + LEAF - 1) / LEAF) * LEAF * LEAF;
    rows_off.push_back(off);
    off += (long)(n - start[c]) * ld;
    off = round_up(off, 16);
  }
  *total = off;
}

// One block column of the distributed factorisation (every rank runs the same program).
int dist_step(munk_dist* h, int c) {
  Workspace* ws = h->ws;
  const int n = h->n, nblk = h->nblk;
  cudaStream_t side = ws->side, side2 = ws->side2, main = h->main;
  cudaEvent_t* diag_ev = ws->events.data();
  cudaEvent_t* top_ev = diag_ev + nblk;      // piece A(c) complete on this rank
  cudaEvent_t* lan_ev = top_ev + nblk;
  cudaEvent_t* panel_ev = lan_ev + nblk;     // piece B(c) complete on this rank
  cudaEvent_t* upd_ev = panel_ev + nblk;     // T(c) done on this rank
  cudaEvent_t* upd1_ev = upd_ev + nblk;      // T(c) done for the first (lowest) block column this rank owns
  cudaEvent_t* top2_ev = upd1_ev + nblk;     // panel c-2 applied to block column c (depth-2 look-ahead, owner of c)
  const bool multi = h->world > 1;
  auto rec = [&](cudaStream_t st, int scratch, int col) {
    Rec r{ws, st, h->linv(col), scratch};
    r.stair_base = h->start[col];
    return r;
  };
  // A[r0:r1, block column col] -= P_col-1[r0:r1] P_col-1[D(col)]^T
  auto la = [&](cudaStream_t st, int scratch, int r0, int r1, int col) -> int {
    if (r1 <= r0) return MUNK_OK;
    Rec r = rec(st, scratch, col);
    const int k = h->w(col - 1);
    const double* X = h->rows(col - 1, r0);
    const double* Y = h->rows(col - 1, h->s(col));
    if (r0 == h->s(col))
      return syrk_update(r, h->rows(col, r0), h->ldp[col], r1 - r0, h->w(col), X, Y, h->ldp[col - 1], k);
    // X/Y live in panel col-1 (its leading dimension), C in panel col
    GemmDesc g{};
    g.M = r1 - r0; g.N = h->w(col); g.K = k;
    g.A = X; g.lda = h->ldp[col - 1]; g.a_layout = OP_KC;
    g.B = Y; g.ldb = h->ldp[col - 1]; g.b_layout = OP_KC;
    g.C = h->rows(col, r0); g.ldc = h->ldp[col];
    g.alpha = -1.0; g.beta = 1.0; g.flags = 0;
    return gemm_launch(ws, g, st, scratch);
  };
  auto ts = [&](cudaStream_t st, int scratch, int r0, int r1, int col) -> int {
    if (r1 <= r0) return MUNK_OK;
    static const int strip_rows = env_int("MUNK_STRIP_ROWS", 1024);
    if (r1 - r0 <= strip_rows) {            // the chain's TS_N: one strip-solve launch
      MUNK_TRY(launch_trsm_strip(h->rows(col, r0), h->ldp[col], r1 - r0, h->rows(col, h->s(col)), h->ldp[col],
                                 h->w(col), h->linv(col), st));
      ws->launches++;
      ws->flops += (double)(r1 - r0) * h->w(col) * h->w(col);
      return MUNK_OK;
    }
    Rec r = rec(st, scratch, col);
    return trsm_rec(r, h->rows(col, r0), h->ldp[col], r1 - r0, h->rows(col, h->s(col)), h->ldp[col], h->w(col), 0);
  };

  // ---- chain stream
  h->tick(side);                                                     // 0: chain stream reaches column c
  if (h->own(c)) {
    // panel c-2 reaches block column c through the panel stream (below), not through the FIFO
    // of this rank's trailing updates
    if (c >= 2) MUNK_CUDA_TRY(cudaStreamWaitEvent(side, top2_ev[c], 0));
    if (c >= 1) MUNK_TRY(la(side, 1, h->s(c), h->s(c + 1), c));
    h->tick(side);                                                   // 1: LA_D done
    Rec r = rec(side, 1, c);
    MUNK_TRY(potrf_diag(r, h->rows(c, h->s(c)), h->ldp[c], h->w(c), 0));
    MUNK_CUDA_TRY(cudaEventRecord(diag_ev[c], side));
    h->tick(side);                                                   // 2: F done
    if (c >= 1) MUNK_CUDA_TRY(cudaStreamWaitEvent(side, lan_ev[c], 0));
    MUNK_TRY(ts(side, 1, h->s(c + 1), h->s(c + 2), c));
  } else {
    h->tick(side);
    h->tick(side);
  }
  h->tick(side);                                                     // 3: TS_N done (owner) / nothing
  if (multi) {
    // piece A(c): leaf inverses + rows D(c) and N(c), contiguous in the panel buffer
    const size_t count = (size_t)h->leaf_elems(c) + (size_t)(h->s(c + 2) - h->s(c)) * h->ldp[c];
    MUNK_NCCL_TRY(nccl().Broadcast(h->linv(c), h->linv(c), count, ncclDouble, owner_of(c, h->world),
                                   h->comm->comm, side));
  }
  MUNK_CUDA_TRY(cudaEventRecord(top_ev[c], side));
  h->tick(side);                                                     // 4: piece A here
  // ---- panel stream
  if (h->own(c)) {
    MUNK_CUDA_TRY(cudaStreamWaitEvent(side2, diag_ev[c], 0));
    MUNK_TRY(ts(side2, 2, h->s(c + 2), n, c));
  }
  // (a receiver starts its part of the broadcast only once piece A(c) is here: until then the root
  // is still factoring, and a waiting NCCL kernel would hold SMs for nothing)
  MUNK_CUDA_TRY(cudaStreamWaitEvent(side2, top_ev[c], 0));
  if (multi && h->s(c + 2) < n) {
    // piece B(c) on the bulk communicator: a later piece A never waits behind it
    const size_t count = (size_t)(n - h->s(c + 2)) * h->ldp[c];
    double* p = h->rows(c, h->s(c + 2));
    MUNK_NCCL_TRY(nccl().Broadcast(p, p, count, ncclDouble, owner_of(c, h->world), h->bulk->comm, side2));
  }
  MUNK_CUDA_TRY(cudaEventRecord(panel_ev[c], side2));
  h->tick(side2);                                                    // 5: piece B here
  if (h->own(c + 1)) {
    // block column c+1 already holds panel c-1 (depth-2 look-ahead of step c-1, this stream)
    MUNK_TRY(la(side2, 2, h->s(c + 2), h->s(c + 3), c + 1));
    MUNK_CUDA_TRY(cudaEventRecord(lan_ev[c + 1], side2));
    MUNK_TRY(la(side2, 2, h->s(c + 3), n, c + 1));
  } else if (c + 1 < nblk) {
    MUNK_CUDA_TRY(cudaEventRecord(lan_ev[c + 1], side2));      // keeps the event defined on every rank
  }
  // ---- depth-2 look-ahead: the owner of block column c+2 applies panel c to it on the panel
  // stream, the moment piece B(c) is here, instead of queueing it behind the trailing updates of
  // its other block columns (measured at 8 GPUs: the next-but-one owner's chain waited 1-2 ms per
  // block column for exactly this update, profiles/r02_dist_trace_n8_fused.log)
  if (h->own(c + 2)) {
    if (c >= 1) MUNK_CUDA_TRY(cudaStreamWaitEvent(side2, upd1_ev[c - 1], 0));   // T(c-1) has left column c+2
    Rec r = rec(side2, 2, c + 2);
    const double* X = h->rows(c, h->s(c + 2));
    MUNK_TRY(syrk_update(r, h->rows(c + 2, h->s(c + 2)), h->ldp[c + 2], n - h->s(c + 2), h->w(c + 2), X, X, h->ldp[c],
                         h->w(c)));
    MUNK_CUDA_TRY(cudaEventRecord(top2_ev[c + 2], side2));
  }
  // ---- caller's stream: panel c applied to this rank's block columns >= c + 3
  MUNK_CUDA_TRY(cudaStreamWaitEvent(main, panel_ev[c], 0));
  bool first = true;
  for (int cc = c + 3; cc < nblk; ++cc) {
    if (!h->own(cc)) continue;
    Rec r = rec(main, 0, cc);
    const double* X = h->rows(c, h->s(cc));
    MUNK_TRY(syrk_update(r, h->rows(cc, h->s(cc)), h->ldp[cc], n - h->s(cc), h->w(cc), X, X, h->ldp[c], h->w(c)));
    if (first) MUNK_CUDA_TRY(cudaEventRecord(upd1_ev[c], main));
    first = false;
  }
  if (first) MUNK_CUDA_TRY(cudaEventRecord(upd1_ev[c], main));
  MUNK_CUDA_TRY(cudaEventRecord(upd_ev[c], main));
  // ---- fused forward substitution with panel c
  if (h->X) {
    cudaStream_t sv = h->solve;
    MUNK_CUDA_TRY(cudaStreamWaitEvent(sv, top_ev[c], 0));
    Rec r = rec(sv, 3, c);
    r.stair = h->stair;
    MUNK_TRY(trsm_left_rec(r, h->rows(c, h->s(c)), h->ldp[c], h->w(c), h->X + (long)h->s(c) * h->ldx, h->ldx,
                           h->ncols, 0, false));
    if (h->s(c + 1) < n) {
      MUNK_CUDA_TRY(cudaStreamWaitEvent(sv, panel_ev[c], 0));
      GemmDesc g{};
      g.M = n - h->s(c + 1); g.N = h->ncols; g.K = h->w(c);
      g.A = h->rows(c, h->s(c + 1)); g.lda = h->ldp[c]; g.a_layout = OP_KC;
      g.B = h->X + (long)h->s(c) * h->ldx; g.ldb = h->ldx; g.b_layout = OP_MC;
      g.C = h->X + (long)h->s(c + 1) * h->ldx; g.ldc = h->ldx;
      g.alpha = -1.0; g.beta = 1.0; g.flags = 0;
      if (h->stair) { g.stair_n = h->stair; g.stair_row0 = h->s(c + 1); g.stair_k0 = h->s(c); }
      MUNK_TRY(gemm_launch(ws, g, sv, 3));
    }
  }
  return MUNK_OK;
}

}  // namespace

extern "C" {

int munk_dist_block(void) { return NB; }

int munk_comm_unique_id(void* id128) {
  if (!id128) return -1;
  if (!nccl().ok) {
    set_last_error("libnccl.so.2 could not be loaded");
    return MUNK_ERR_NCCL;
  }
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  MUNK_NCCL_TRY(nccl().GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return MUNK_OK;
}

int munk_comm_create(int device, int rank, int world, const void* id128, int max_ctas, munk_comm_t** out) {
  if (!out) return -5;
  *out = nullptr;
  if (world < 1 || rank < 0 || rank >= world) return -2;
  munk_comm* c = new (std::nothrow) munk_comm();
  if (!c) return -5;
  c->rank = rank; c->world = world; c->device = device;
  if (world > 1) {
    if (!id128) { delete c; return -4; }
    if (!nccl().ok) {
      delete c;
      set_last_error("libnccl.so.2 could not be loaded");
      return MUNK_ERR_NCCL;
    }
    DeviceGuard guard(device);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    // A collective's kernel occupies one SM per CTA from the moment it is launched until its data
    // has arrived - on the receivers that is the whole time the root still computes - and an SM
    // that holds an NCCL CTA cannot hold a GEMM tile.  The panel pieces are small (MBs): a few
    // CTAs move them at link speed, so the communicator is capped (max_ctas > 0).
    ncclResult_t r = ncclInvalidUsage;
    if (max_ctas > 0 && nccl().CommInitRankConfig) {
      ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
      cfg.minCTAs = 1;
      cfg.maxCTAs = max_ctas;
      r = nccl().CommInitRankConfig(&c->comm, world, id, rank, &cfg);
      if (r != ncclSuccess) c->comm = nullptr;
    }
    if (r != ncclSuccess) r = nccl().CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
      set_last_error("ncclCommInitRank failed: %s", nccl().GetErrorString(r));
      delete c;
      return MUNK_ERR_NCCL + (int)r;
    }
  }
  *out = c;
  return MUNK_OK;
}

int munk_comm_destroy(munk_comm_t* c) {
  if (!c) return MUNK_OK;
  if (c->comm) {
    DeviceGuard guard(c->device);
    nccl().CommDestroy(c->comm);
  }
  delete c;
  return MUNK_OK;
}

size_t munk_dist_storage_elems(int n) {
  std::vector<int> start, ldp;
  std::vector<long> lo, ro;
  long total = 0;
  layout(std::max(n, 0), start, ldp, lo, ro, &total);
  return (size_t)total;
}

int munk_dist_create(munk_workspace_t* wsp, munk_comm_t* comm, munk_comm_t* bulk, int n, double* storage,
                     munk_dist_t** out) {
  Workspace* ws = reinterpret_cast<Workspace*>(wsp);
  if (!ws) return -1;
  if (n <= 0) return -3;
  if (!storage || ((uintptr_t)storage & 15)) return -4;
  if (!out) return -5;
  *out = nullptr;
  DeviceGuard guard(ws->device);
  munk_dist* h = new (std::nothrow) munk_dist();
  if (!h) return -5;
  h->ws = ws; h->comm = comm; h->bulk = bulk ? bulk : comm;
  h->rank = comm ? comm->rank : 0;
  h->world = comm ? comm->world : 1;
  h->n = n; h->storage = storage;
  long total = 0;
  layout(n, h->start, h->ldp, h->linv_off, h->rows_off, &total);
  h->nblk = (int)h->start.size() - 1;
  // device copies of the layout
  std::vector<int> colblk(n);
  std::vector<unsigned char> owned(h->nblk);
  for (int c = 0; c < h->nblk; ++c) {
    owned[c] = h->own(c) ? 1 : 0;
    for (int j = h->start[c]; j < h->start[c + 1]; ++j) colblk[j] = c;
  }
  cudaError_t e = cudaMalloc((void**)&h->d_colblk, (size_t)n * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_start, (size_t)(h->nblk + 1) * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_ldp, (size_t)h->nblk * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_rows_off, (size_t)h->nblk * sizeof(long));
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_owned, (size_t)h->nblk);
  if (e == cudaSuccess) e = cudaMemcpy(h->d_colblk, colblk.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->d_start, h->start.data(), (size_t)(h->nblk + 1) * sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->d_ldp, h->ldp.data(), (size_t)h->nblk * sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->d_rows_off, h->rows_off.data(), (size_t)h->nblk * sizeof(long), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->d_owned, owned.data(), (size_t)h->nblk, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->solve, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->solve_done, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    set_last_error("munk_dist_create: %s", cudaGetErrorString(e));
    munk_dist_destroy(h);
    return MUNK_ERR_CUDA + (int)e;
  }
  *out = h;
  return MUNK_OK;
}

int munk_dist_destroy(munk_dist_t* h) {
  if (!h) return MUNK_OK;
  DeviceGuard guard(h->ws->device);
  cudaFree(h->d_colblk); cudaFree(h->d_start); cudaFree(h->d_ldp); cudaFree(h->d_rows_off); cudaFree(h->d_owned);
  if (h->solve) cudaStreamDestroy(h->solve);
  if (h->solve_done) cudaEventDestroy(h->solve_done);
  delete h;
  return MUNK_OK;
}

int munk_dist_blocks(munk_dist_t* h) { return h ? h->nblk : -1; }

int munk_dist_assemble(munk_dist_t* h, int nnz, const int* u, const int* v, const double* lval,
                       const double* ldiag, double lam, void* stream_) {
  if (!h) return -1;
  if (nnz < 0) return -2;
  if (nnz > 0 && (!u || !v)) return -3;
  if ((lval == nullptr) != (ldiag == nullptr)) return -5;
  Workspace* ws = h->ws;
  cudaStream_t stream = (cudaStream_t)stream_;
  DeviceGuard guard(ws->device);
  int* deg = nullptr;
  if (!ldiag) {
    MUNK_TRY(ws->ensure_splitk(0, std::max(SPLITK_RESERVE, (size_t)(h->n + 2) * sizeof(int))));
    deg = (int*)ws->splitk[0];
    MUNK_CUDA_TRY(cudaMemsetAsync(deg, 0, (size_t)h->n * sizeof(int), stream));
    if (nnz > 0) {
      degree_count_kernel<<<(nnz + 255) / 256, 256, 0, stream>>>(h->n, nnz, u, v, deg, ws->info + 1);
      MUNK_CUDA_TRY(cudaGetLastError());
    }
  }
  for (int c = 0; c < h->nblk; ++c) {
    if (!h->own(c)) continue;
    const int rows = h->n - h->start[c];
    dim3 grid(rows, std::min(4, (h->ldp[c] / 2 + 255) / 256));
    fill_panel_kernel<<<grid, 256, 0, stream>>>(rows, h->w(c), h->ldp[c], h->start[c], lam, deg, ldiag,
                                                h->storage + h->rows_off[c]);
    MUNK_CUDA_TRY(cudaGetLastError());
    ws->launches++;
  }
  if (nnz > 0) {
    scatter_panel_kernel<<<(nnz + 255) / 256, 256, 0, stream>>>(h->n, nnz, u, v, lval, lam, h->d_colblk, h->d_start,
                                                                h->d_rows_off, h->d_ldp, h->d_owned, h->storage,
                                                                ws->info + 1);
    MUNK_CUDA_TRY(cudaGetLastError());
    ws->launches += 2;
  }
  return MUNK_OK;
}

int munk_dist_set_rhs(munk_dist_t* h, double* X, long ldx, int ncols, const int* stair) {
  if (!h) return -1;
  if (X && (ldx < ncols || (ldx & 1) || ((uintptr_t)X & 15))) return -3;
  if (ncols < 0) return -4;
  if (h->running) return -1;
  h->X = (X && ncols > 0) ? X : nullptr;
  h->ldx = ldx; h->ncols = ncols; h->stair = stair;
  return MUNK_OK;
}

int munk_dist_potrf_begin(munk_dist_t* h, void* stream_) {
  if (!h) return -1;
  if (h->running) return -1;
  Workspace* ws = h->ws;
  DeviceGuard guard(ws->device);
  h->main = (cudaStream_t)stream_;
  MUNK_TRY(ws->ensure_side(7 * (size_t)h->nblk + 2));
  for (int i = 0; i < 4; ++i) MUNK_TRY(ws->ensure_splitk(i, SPLITK_RESERVE));
  cudaEvent_t fork_ev = ws->events[7 * h->nblk];
  MUNK_CUDA_TRY(cudaEventRecord(fork_ev, h->main));
  MUNK_CUDA_TRY(cudaStreamWaitEvent(ws->side, fork_ev, 0));
  MUNK_CUDA_TRY(cudaStreamWaitEvent(ws->side2, fork_ev, 0));
  MUNK_CUDA_TRY(cudaStreamWaitEvent(h->solve, fork_ev, 0));
  h->next = 0;
  h->running = true;
  h->trace = getenv("MUNK_TRACE_DIST") != nullptr;
  h->tr.clear();
  return MUNK_OK;
}

/* Enqueues block column `next`; *remaining receives the number of block columns still to go. */
int munk_dist_potrf_step(munk_dist_t* h, int* remaining) {
  if (!h || !h->running) return -1;
  DeviceGuard guard(h->ws->device);
  if (h->next < h->nblk) {
    MUNK_TRY(dist_step(h, h->next));
    h->next++;
  }
  if (remaining) *remaining = h->nblk - h->next;
  return MUNK_OK;
}

int munk_dist_potrf_end(munk_dist_t* h) {
  if (!h || !h->running) return -1;
  Workspace* ws = h->ws;
  DeviceGuard guard(ws->device);
  while (h->next < h->nblk) {
    MUNK_TRY(dist_step(h, h->next));
    h->next++;
  }
  cudaEvent_t join_ev = ws->events[7 * h->nblk + 1];
  MUNK_CUDA_TRY(cudaEventRecord(join_ev, ws->side));
  MUNK_CUDA_TRY(cudaStreamWaitEvent(h->main, join_ev, 0));
  cudaEvent_t* panel_ev = ws->events.data() + 3 * h->nblk;
  MUNK_CUDA_TRY(cudaStreamWaitEvent(h->main, panel_ev[h->nblk - 1], 0));
  MUNK_CUDA_TRY(cudaEventRecord(h->solve_done, h->solve));
  MUNK_CUDA_TRY(cudaStreamWaitEvent(h->main, h->solve_done, 0));
  // side2 may still hold the (empty) tail after panel_ev: one more join through an event on it
  MUNK_CUDA_TRY(cudaEventRecord(join_ev, ws->side2));
  MUNK_CUDA_TRY(cudaStreamWaitEvent(h->main, join_ev, 0));
  h->running = false;
  if (h->trace) {
    cudaStreamSynchronize(h->main);
    fprintf(stderr, "dist potrf n=%d rank %d/%d: col owner | reach, LA_D, F, lan+TS_N, bcastA done, pieceB done (ms since start)\n",
            h->n, h->rank, h->world);
    for (int c = 0; c < h->nblk; ++c) {
      float t[6];
      for (int i = 0; i < 6; ++i) cudaEventElapsedTime(&t[i], h->tr[0], h->tr[6 * c + i]);
      fprintf(stderr, "  %3d %d | %8.3f %8.3f %8.3f %8.3f %8.3f %8.3f\n", c, owner_of(c, h->world), t[0], t[1], t[2], t[3],
              t[4], t[5]);
    }
    for (cudaEvent_t e : h->tr) cudaEventDestroy(e);
    h->tr.clear();
  }
  return MUNK_OK;
}

int munk_dist_solve_backward(munk_dist_t* h, double* X, long ldx, int ncols, void* stream_) {
  if (!h) return -1;
  if (!X || ldx < ncols || (ldx & 1) || ((uintptr_t)X & 15)) return -2;
  if (ncols <= 0) return MUNK_OK;
  if (h->running) return -1;
  Workspace* ws = h->ws;
  cudaStream_t stream = (cudaStream_t)stream_;
  DeviceGuard guard(ws->device);
  MUNK_TRY(ws->ensure_splitk(0, SPLITK_RESERVE));
  const int n = h->n;
  for (int c = h->nblk - 1; c >= 0; --c) {
    double* Xc = X + (long)h->start[c] * ldx;
    if (h->s(c + 1) < n) {
      // X[D(c)] -= P_c[below]^T X[below]
      GemmDesc g{};
      g.M = h->w(c); g.N = ncols; g.K = n - h->s(c + 1);
      g.A = h->rows(c, h->s(c + 1)); g.lda = h->ldp[c]; g.a_layout = OP_MC;
      g.B = X + (long)h->s(c + 1) * ldx; g.ldb = ldx; g.b_layout = OP_MC;
      g.C = Xc; g.ldc = ldx;
      g.alpha = -1.0; g.beta = 1.0; g.flags = 0;
      MUNK_TRY(gemm_launch(ws, g, stream, 0));
    }
    Rec r{ws, stream, h->linv(c), 0};
    MUNK_TRY(trsm_left_rec(r, h->rows(c, h->s(c)), h->ldp[c], h->w(c), Xc, ldx, ncols, 0, true));
  }
  return MUNK_OK;
}

int munk_dist_unpack(munk_dist_t* h, double* L, long ld, void* stream_) {
  if (!h) return -1;
  if (!L || ld < h->n) return -2;
  cudaStream_t stream = (cudaStream_t)stream_;
  DeviceGuard guard(h->ws->device);
  for (int c = 0; c < h->nblk; ++c) {
    const int rows = h->n - h->start[c];
    unpack_panel_kernel<<<rows, 128, 0, stream>>>(rows, h->w(c), h->ldp[c], h->storage + h->rows_off[c],
                                                  L + (long)h->start[c] * ld + h->start[c], ld);
    MUNK_CUDA_TRY(cudaGetLastError());
  }
  return MUNK_OK;
}

}  // extern "C"

// FP64 tensor-core GEMM for sm_100a:  C = alpha * A * B^T + beta * C   (row-major C).
//
// B200 has no tcgen05 kind for f64, so the FP64 tensor path is the warp-level DMMA
// (mma.sync.m8n8k4.f64).  The kernel is built around it:
//   * one CTA = 128x128 output tile, 8 warps (2x4, each 64x32 = 8x4 DMMA tiles, 64 accumulator
//     doubles per lane); lane 0 of the warps doubles as the TMA producer (no ninth warp: it
//     would cap the kernel at 168 registers and spill into the main loop);
//   * operands are staged by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) into a 6-stage
//     ring of 128x16 (A) + 128x16 (B) tiles guarded by full/empty mbarriers, 4 stages in flight;
//   * either operand may be stored k-contiguous ("KC", rows of 16 doubles = 128 B) or
//     free-index-contiguous ("MC", eight 16x16 boxes); one k-permutation inside the
//     16-deep stage (k(j,s) below) makes the 64-bit fragment loads bank-conflict-free
//     for both layouts under the 128-byte swizzle;
//   * structure flags skip the k-range where a triangular operand is zero and the tiles
//     above the diagonal of a symmetric result, so triangular products cost n^3/3.
// Used for every dense contraction on the MUNK hot path (reference call sites:
// munk/munk/munk.py:83 inv, :88 v.dot(diag), :109-110 pinv(..).dot(..);
// scripts/compute_embeddings.py:81 np.dot(source_X, target_X.T)).
#include "dgemm.cuh"
#include "workspace.cuh"

#include <algorithm>
#include <cstdarg>

namespace munk {

namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int STAGES = 6;
constexpr int CONSUMER_WARPS = 8;
constexpr int GEMM_THREADS = CONSUMER_WARPS * 32;   // thread 0 doubles as the TMA producer
constexpr int LOOKAHEAD = 4;                        // stages in flight ahead of the consumers
constexpr int TILE_BYTES = BM * BK * 8;         // 16 KB
constexpr int STAGE_BYTES = 2 * TILE_BYTES;     // A + B
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 128 /*barriers*/;

struct KernelArgs {
  int M, N, K;
  double* C;
  long ldc;
  double alpha, beta;
  int flags;
  int nsplit;
  double* ws;       // split-K partials, slice z at ws + z * ws_slice, row stride ws_ld
  long ws_slice;
  long ws_ld;
  int mt, nt, gn;   // tile grid and rasterisation strip width (in tiles)
  const int* stair_n;   // see GemmDesc
  const int* stair_m;
  int stair_row0, stair_k0;
};

// k index (0..15 inside a stage) read by lane-quad position j = lane & 3 at sub-step s:
//   k(j,s) = 8*(j>>1) + (j&1) + 2*((j+s)&3)
// (a bijection of {0..3}x{0..3} onto 0..15; A and B use the same map, so the contraction
// is merely re-ordered).  It is chosen so that, with SWIZZLE_128B, the 16 lanes of a
// half-warp hit 16 distinct 8-byte bank pairs for both operand layouts.
__device__ __forceinline__ int kmap(int j, int s) { return 8 * (j >> 1) + (j & 1) + 2 * ((j + s) & 3); }

// Byte offset (inside a 16 KB operand tile) of element (x = wbase + r8 [+ 8t], k(j,s)).
template <bool KC>
__device__ __forceinline__ uint32_t frag_base(int wbase, int r8, int j, int s) {
  const int k = kmap(j, s);
  if (KC) {
    // tile = 128 rows x 128 B; 16-byte chunk index XOR (row & 7).  (wbase % 8 == 0)
    return (uint32_t)((wbase + r8) * 128 + ((((k >> 1) ^ r8) & 7) << 4) + (k & 1) * 8);
  } else {
    // tile = 8 boxes (16 free-index values each) x 16 k-rows x 128 B; chunk XOR (k & 7).
    // (wbase % 16 == 0; the 8t part is added by frag_addr.)
    return (uint32_t)((wbase >> 4) * 2048 + k * 128 + ((((r8 >> 1) ^ (k & 7)) & 7) << 4) +
                      (r8 & 1) * 8);
  }
}
template <bool KC>
__device__ __forceinline__ uint32_t frag_addr(uint32_t base, int t) {
  if (KC) return base + t * 1024;                       // 8 rows further
  return (base ^ ((t & 1) << 6)) + (t >> 1) * 2048;      // +8 inside the box flips chunk bit 2
}

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
dgemm_dmma_tma_kernel(const __grid_constant__ CUtensorMap tmA,
                      const __grid_constant__ CUtensorMap tmB, const KernelArgs g) {
  // Tile rasterisation: strips of `gn` tile columns, tile rows fastest inside a strip, so a wave
  // of 148 CTAs covers about 12 x 12 tiles and shares 12 + 12 operand panels through L2 instead
  // of 148 + 1 (the m-fastest order re-streamed all of A once per tile column: 210 GB of DRAM
  // reads for the C3 score product, profiles/r01_ncu_summary.md).
  int tile_m, tile_n;
  {
    const int per = g.gn * g.mt;
    const int strip = (int)blockIdx.x / per;
    const int rem = (int)blockIdx.x - strip * per;
    const int width = min(g.gn, g.nt - strip * g.gn);
    tile_m = rem / width;
    tile_n = strip * g.gn + (rem - tile_m * width);
  }
  const int m0 = tile_m * BM;
  const int n0 = tile_n * BN;
  if ((g.flags & GF_TILE_LOWER) && n0 >= m0 + BM) return;

  // k-range of this tile, in 16-deep iterations.
  int klo = 0, khi = g.K;
  if (g.flags & GF_KLO_M) klo = max(klo, m0);
  if (g.flags & GF_KLO_N) klo = max(klo, n0);
  if (g.flags & GF_KHI_M) khi = min(khi, m0 + BM);
  if (g.flags & GF_KHI_N) khi = min(khi, n0 + BN);
  if (g.stair_n) {
    const int first = g.stair_n[tile_n];
    if (g.stair_row0 + m0 + BM <= first) return;        // whole tile above the staircase: stays zero
    klo = max(klo, first - g.stair_k0);
  }
  if (g.stair_m) klo = max(klo, g.stair_m[tile_m] - g.stair_k0);
  int it_lo = klo / BK;
  int it_hi = (khi + BK - 1) / BK;
  if (it_hi < it_lo) it_hi = it_lo;
  if (g.nsplit > 1) {
    const int nit = it_hi - it_lo;
    const int per = (nit + g.nsplit - 1) / g.nsplit;
    it_lo = min(it_hi, it_lo + (int)blockIdx.z * per);
    it_hi = min(it_hi, it_lo + per);
  }

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;
  uint8_t* tiles = smem_raw + pad;                    // 1024-byte aligned (swizzle atom)
  const uint32_t tiles_u32 = raw + pad;
  const uint32_t bars = tiles_u32 + STAGES * STAGE_BYTES;   // full[STAGES], empty[STAGES]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bars + 8 * s, 1);                          // full: producer's expect_tx arrive
      mbar_init(bars + 8 * (STAGES + s), CONSUMER_WARPS);  // empty: one arrive per consumer warp
    }
    mbar_fence_init();
  }
  __syncthreads();

  // ---------------- producer role: thread 0 issues all TMA traffic ----------------
  // There is no dedicated producer warp: a ninth warp would put three warps on one scheduler
  // and cap every thread at 168 registers (64 accumulator doubles + fragments need ~180; the
  // 9-warp build spilled into the main loop).  Thread 0 refills, at the top of iteration `it`, the
  // stage of iteration it + LOOKAHEAD; that slot was consumed STAGES - LOOKAHEAD = 2 iterations
  // ago, so the empty-barrier wait is satisfied unless a warp lags two iterations behind.
  int pstage = 0;                 // producer ring position
  uint32_t pphase = 0;
  // Called by lane 0 of every warp.  A free-index-contiguous ("MC") operand tile is eight 16x16
  // boxes: warp w issues box w, so no single warp carries all the copy instructions (one thread
  // issuing 16 copies per iteration slowed its warp, and with it the ring, by ~10 %); a
  // k-contiguous tile is one box, issued by warp 0.  Thread 0 posts the transaction count; a copy
  // that completes before the count is posted just drives it negative for a moment, the phase
  // cannot complete before thread 0's arrival.
  auto issue = [&](int it) {
    const bool work = (warp == 0) || !A_KC || !B_KC;
    if (work) {
      mbar_wait(bars + 8 * (STAGES + pstage), pphase ^ 1);
      const uint32_t full = bars + 8 * pstage;
      if (warp == 0) mbar_arrive_expect_tx(full, STAGE_BYTES);
      const uint32_t sa = tiles_u32 + pstage * STAGE_BYTES;
      const uint32_t sb = sa + TILE_BYTES;
      const int k0 = it * BK;
      if (A_KC) {
        if (warp == 0) tma_load_2d(sa, &tmA, k0, m0, full);
      } else {
        tma_load_2d(sa + warp * 2048, &tmA, m0 + 16 * warp, k0, full);
      }
      if (B_KC) {
        if (warp == 0) tma_load_2d(sb, &tmB, k0, n0, full);
      } else {
        tma_load_2d(sb + warp * 2048, &tmB, n0 + 16 * warp, k0, full);
      }
    }
    if (++pstage == STAGES) { pstage = 0; pphase ^= 1; }
  };
  if (lane == 0) {
    if (warp == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
    }
    for (int it = it_lo; it < it_hi && it < it_lo + LOOKAHEAD; ++it) issue(it);
  }
  if (g.beta != 0.0 && g.nsplit == 1) {
    // Pull this tile of C into L2 while the main loop runs.  Without it the CTAs of a wave (which
    // run in lockstep) all miss to DRAM in their epilogues at the same moment: ~19 MB in and
    // ~19 MB out in one burst, 8.5 us per tile with the tensor pipe idle (tools/syrk_probe.py:
    // K = 512 update at 28 TFLOP/s with beta = 1 vs 31 with beta = 0).
    for (int idx = threadIdx.x; idx < BM * (BN / 16); idx += GEMM_THREADS) {
      const int r = m0 + idx / (BN / 16), c = n0 + (idx % (BN / 16)) * 16;
      if (r < g.M && c < g.N) prefetch_l2(g.C + (long)r * g.ldc + c);
    }
  }

  // ---------------- all 8 warps: DMMA over the staged tiles ----------------
  const int wm = (warp >> 2) * 64;   // warp tile origin inside the CTA tile
  const int wn = (warp & 3) * 32;
  const int r8 = lane >> 2;
  const int j = lane & 3;

  uint32_t offA[4], offB[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    offA[s] = frag_base<A_KC>(wm, r8, j, s);
    offB[s] = TILE_BYTES + frag_base<B_KC>(wn, r8, j, s);
  }

  double acc[8][4][2];
#pragma unroll
  for (int t = 0; t < 8; ++t)
#pragma unroll
    for (int u = 0; u < 4; ++u) acc[t][u][0] = acc[t][u][1] = 0.0;

  {
    int stage = 0;
    uint32_t phase = 0;
    for (int it = it_lo; it < it_hi; ++it) {
      if (lane == 0 && it + LOOKAHEAD < it_hi) issue(it + LOOKAHEAD);
      __syncwarp();
      mbar_wait(bars + 8 * stage, phase);
      const uint8_t* st = tiles + stage * STAGE_BYTES;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        double a[8], b[4];
#pragma unroll
        for (int t = 0; t < 8; ++t)
          a[t] = *reinterpret_cast<const double*>(st + frag_addr<A_KC>(offA[s], t));
#pragma unroll
        for (int u = 0; u < 4; ++u)
          b[u] = *reinterpret_cast<const double*>(st + frag_addr<B_KC>(offB[s], u));
#pragma unroll
        for (int t = 0; t < 8; ++t)
#pragma unroll
          for (int u = 0; u < 4; ++u) dmma884(acc[t][u][0], acc[t][u][1], a[t], b[u]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 8 * (STAGES + stage));
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  }

  // ---------------- epilogue ----------------
  const bool split = g.nsplit > 1;
  double* __restrict__ out = split ? g.ws + (long)blockIdx.z * g.ws_slice : g.C;
  const long ldo = split ? g.ws_ld : g.ldc;
  const double alpha = split ? 1.0 : g.alpha;
  const double beta = split ? 0.0 : g.beta;
  const bool mirror = !split && (g.flags & GF_SYM_MIRROR) && (n0 + BN <= m0);

  // beta != 0: bring the old C tile into the (now idle) pipeline shared memory with cp.async —
  // 32 independent 16-byte copies per thread, one round trip to L2 (the producer warp prefetched
  // the tile there) — instead of 32 dependent load -> fma -> store round trips per lane, which is
  // what the compiler emits for a read-modify-write through one pointer (~10 us per tile).
  // Row stride 1088 B keeps the 128-bit fragment reads below bank-conflict-free.
  constexpr int CS = BN * 8 + 64;
  if (beta != 0.0) {
    asm volatile("bar.sync 1, 256;" ::: "memory");      // every consumer warp is done with the ring
    const int tid = threadIdx.x;                         // 0..255 (consumer warps only)
#pragma unroll 4
    for (int i = 0; i < (BM * BN / 2) / 256; ++i) {
      const int q = i * 256 + tid;                       // 16-byte chunk of the tile
      const int rl = q >> 6, cl = (q & 63) * 2;
      const int row = m0 + rl, col = n0 + cl;
      const uint32_t dst = tiles_u32 + rl * CS + cl * 8;
      const double* src = out + (long)row * ldo + col;
      if (row < g.M && col + 1 < g.N)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
      else if (row < g.M && col < g.N)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    asm volatile("bar.sync 1, 256;" ::: "memory");
  }

#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const int row = m0 + wm + 8 * t + r8;
    if (row >= g.M) continue;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int col = n0 + wn + 8 * u + 2 * j;
      if (col >= g.N) continue;
      double v0 = alpha * acc[t][u][0];
      double v1 = alpha * acc[t][u][1];
      double* p = out + (long)row * ldo + col;
      const double* sp = reinterpret_cast<const double*>(tiles + (wm + 8 * t + r8) * CS + (wn + 8 * u + 2 * j) * 8);
      if (col + 1 < g.N) {
        if (beta != 0.0) {
          const double2 old = *reinterpret_cast<const double2*>(sp);
          v0 += beta * old.x;
          v1 += beta * old.y;
        }
        *reinterpret_cast<double2*>(p) = make_double2(v0, v1);
        if (mirror) {
          g.C[(long)col * g.ldc + row] = v0;
          g.C[(long)(col + 1) * g.ldc + row] = v1;
        }
      } else {
        if (beta != 0.0) v0 += beta * sp[0];
        p[0] = v0;
        if (mirror) g.C[(long)col * g.ldc + row] = v0;
      }
    }
  }
}

// C = alpha * sum_z ws[z] + beta * C over the tiles the GEMM produced.
__global__ void splitk_reduce_kernel(const KernelArgs g) {
  const int col = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  const int row = blockIdx.y;
  if (col >= g.N) return;
  if ((g.flags & GF_TILE_LOWER) && (col / BN) * BN >= (row / BM) * BM + BM) return;
  if (g.stair_n && g.stair_row0 + (row / BM) * BM + BM <= g.stair_n[col / BN]) return;
  const bool pair = col + 1 < g.N;
  double s0 = 0.0, s1 = 0.0;
  for (int z = 0; z < g.nsplit; ++z) {
    const double* p = g.ws + (long)z * g.ws_slice + (long)row * g.ws_ld + col;
    s0 += p[0];
    if (pair) s1 += p[1];
  }
  double* c = g.C + (long)row * g.ldc + col;
  double v0 = g.alpha * s0, v1 = g.alpha * s1;
  if (g.beta != 0.0) {
    v0 += g.beta * c[0];
    if (pair) v1 += g.beta * c[1];
  }
  c[0] = v0;
  if (pair) c[1] = v1;
}

__global__ void scale_kernel(double* C, long ldc, int M, int N, double beta) {
  const int col = blockIdx.y * blockDim.x + threadIdx.x;
  const int row = blockIdx.x;
  if (col < N && row < M) C[(long)row * ldc + col] = beta == 0.0 ? 0.0 : beta * C[(long)row * ldc + col];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess)
      p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// Tensor map of one operand view: `rows` free-index values x `K` contraction values.
int make_operand_map(CUtensorMap* tm, const double* base, long ld, int layout, int rows, int K) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) {
    set_last_error("cuTensorMapEncodeTiled entry point not available");
    return MUNK_ERR_TMAP;
  }
  if (((uintptr_t)base & 15) || (ld & 1)) {
    set_last_error("GEMM operand must be 16-byte aligned with an even leading dimension");
    return -1;
  }
  cuuint64_t dims[2];
  cuuint32_t box[2];
  if (layout == OP_KC) {
    dims[0] = (cuuint64_t)K;  dims[1] = (cuuint64_t)rows;
    box[0] = BK;              box[1] = BM;
  } else {
    dims[0] = (cuuint64_t)rows;  dims[1] = (cuuint64_t)K;
    box[0] = 16;                 box[1] = BK;
  }
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%d K=%d ld=%ld)", (int)r,
                   rows, K, ld);
    return MUNK_ERR_TMAP + (int)r;
  }
  return MUNK_OK;
}

template <bool A_KC, bool B_KC>
int launch_variant(const CUtensorMap& ta, const CUtensorMap& tb, const KernelArgs& ka, dim3 grid,
                   cudaStream_t stream) {
  static bool configured[MUNK_MAX_DEVICES] = {};       // the attribute is per device
  auto kern = dgemm_dmma_tma_kernel<A_KC, B_KC>;
  int dev = 0;
  MUNK_CUDA_TRY(cudaGetDevice(&dev));
  if (dev >= MUNK_MAX_DEVICES || !configured[dev]) {
    MUNK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    if (dev < MUNK_MAX_DEVICES) configured[dev] = true;
  }
  kern<<<grid, GEMM_THREADS, SMEM_BYTES, stream>>>(ta, tb, ka);
  MUNK_CUDA_TRY(cudaGetLastError());
  return MUNK_OK;
}

// k-iterations a tile executes (host mirror of the kernel's range logic).
inline int tile_iters(const GemmDesc& g, int m0, int n0) {
  int klo = 0, khi = g.K;
  if (g.flags & GF_KLO_M) klo = std::max(klo, m0);
  if (g.flags & GF_KLO_N) klo = std::max(klo, n0);
  if (g.flags & GF_KHI_M) khi = std::min(khi, m0 + BM);
  if (g.flags & GF_KHI_N) khi = std::min(khi, n0 + BN);
  const int lo = klo / BK, hi = (khi + BK - 1) / BK;
  return std::max(0, hi - lo);
}

}  // namespace

double gemm_flops(const GemmDesc& g) {
  double f = 0.0;
  for (int m0 = 0; m0 < g.M; m0 += BM)
    for (int n0 = 0; n0 < g.N; n0 += BN) {
      if ((g.flags & GF_TILE_LOWER) && n0 >= m0 + BM) continue;
      const double rows = std::min(BM, g.M - m0), cols = std::min(BN, g.N - n0);
      f += 2.0 * rows * cols * (double)tile_iters(g, m0, n0) * BK;
    }
  return f;
}

int gemm_launch(Workspace* ws, const GemmDesc& g, cudaStream_t stream, int scratch) {
  if (g.M <= 0 || g.N <= 0) return MUNK_OK;
  if (((uintptr_t)g.C & 15) || (g.ldc & 1)) {
    set_last_error("GEMM output must be 16-byte aligned with an even leading dimension");
    return -1;
  }
  if (g.K <= 0) {
    dim3 grid(g.M, (g.N + 255) / 256);
    scale_kernel<<<grid, 256, 0, stream>>>(g.C, g.ldc, g.M, g.N, g.beta);
    MUNK_CUDA_TRY(cudaGetLastError());
    if (ws) ws->launches++;
    return MUNK_OK;
  }
  CUtensorMap ta, tb;
  MUNK_TRY(make_operand_map(&ta, g.A, g.lda, g.a_layout, g.M, g.K));
  MUNK_TRY(make_operand_map(&tb, g.B, g.ldb, g.b_layout, g.N, g.K));

  const int mt = (g.M + BM - 1) / BM, nt = (g.N + BN - 1) / BN;
  long active = 0;
  int max_it = 0;
  for (int a = 0; a < mt; ++a)
    for (int b = 0; b < nt; ++b) {
      if ((g.flags & GF_TILE_LOWER) && b * BN >= a * BM + BM) continue;
      ++active;
      max_it = std::max(max_it, tile_iters(g, a * BM, b * BN));
    }

  KernelArgs ka;
  ka.M = g.M; ka.N = g.N; ka.K = g.K;
  ka.C = g.C; ka.ldc = g.ldc;
  ka.alpha = g.alpha; ka.beta = g.beta;
  ka.flags = g.flags;
  ka.stair_n = g.stair_n; ka.stair_m = g.stair_m;
  ka.stair_row0 = g.stair_row0; ka.stair_k0 = g.stair_k0;
  ka.nsplit = 1; ka.ws = nullptr; ka.ws_slice = 0; ka.ws_ld = 0;

  // Split-K whenever the tile grid cannot fill the 148 SMs: one 128x128 tile advances only 16 of
  // k every ~2 us (DMMA rate of one SM), so the small and mid-size products on the critical path
  // of the recursive factorisations are latency-bound unless their k-range is spread out.
  // Each split keeps at least 2 k-iterations; partials are reduced in a fixed order.
  int nsplit = 1;
  static const int split_min_it = [] { const char* e = getenv("MUNK_SPLITK_MIN_IT"); return e ? std::max(2, atoi(e)) : 4; }();
  if (ws && active > 0 && active <= 111 && max_it >= split_min_it && !(g.flags & GF_SYM_MIRROR)) {
    nsplit = (int)std::min<long>((148 + active - 1) / active, max_it / 2);
    nsplit = std::max(1, std::min(nsplit, 74));
  }
  if (nsplit > 1) {
    ka.nsplit = nsplit;
    ka.ws_ld = round_up(g.N, 2);
    ka.ws_slice = (long)g.M * ka.ws_ld;
    MUNK_TRY(ws->ensure_splitk(scratch, std::max(SPLITK_RESERVE, (size_t)nsplit * ka.ws_slice * 8)));
    ka.ws = ws->splitk[scratch];
  }

  static const int strip = [] { const char* e = getenv("MUNK_GEMM_STRIP"); return e ? std::max(1, atoi(e)) : 12; }();
  ka.mt = mt; ka.nt = nt; ka.gn = std::min(nt, strip);
  dim3 grid((unsigned)(mt * nt), 1, nsplit);
  int st;
  if (g.a_layout == OP_KC && g.b_layout == OP_KC) st = launch_variant<true, true>(ta, tb, ka, grid, stream);
  else if (g.a_layout == OP_KC) st = launch_variant<true, false>(ta, tb, ka, grid, stream);
  else if (g.b_layout == OP_KC) st = launch_variant<false, true>(ta, tb, ka, grid, stream);
  else st = launch_variant<false, false>(ta, tb, ka, grid, stream);
  MUNK_TRY(st);
  if (ws) { ws->launches++; ws->flops += gemm_flops(g); }

  if (nsplit > 1) {
    dim3 rgrid((g.N / 2 + 128) / 128, g.M);
    splitk_reduce_kernel<<<rgrid, 128, 0, stream>>>(ka);
    MUNK_CUDA_TRY(cudaGetLastError());
    ws->launches++;
  }
  return MUNK_OK;
}

}  // namespace munk

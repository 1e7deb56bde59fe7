// separate_scores on the device (reference: munk/munk/munk.py:20-68) and the permutation-test
// statistic mean(H_H) - mean(other) (experiments/homology-scores-permutation-test/permtest.py:58-66).
//
// The reference builds six n1 x n2 boolean masks and selects with them; each selection comes out
// in row-major order.  Because the masks are "landmark row", "landmark column" and two sparse
// cell sets (landmark pairs P, homolog pairs HH), the destination of every cell is a closed
// form of per-row offsets (computed by the host from the integer index lists) and the rank of
// its column among landmark / non-landmark columns.  One streaming pass over S (HBM-bound,
// n1*n2*8 bytes read, about the same written) replaces the masks.
#include "../../include/munk_b200.h"
#include "common.cuh"

#include <algorithm>

namespace munk {
namespace {

struct SepArgs {
  int n1, n2;
  const double* S; long lds;
  const unsigned char* rowL;   // [n1] 1 if the row is a landmark row
  const unsigned char* colL;   // [n2]
  const int* colrank;          // [n2] number of landmark columns before column j
  const int* p_ptr;  const int* p_col;   // CSR over rows: landmark-pair cells, columns ascending
  const int* h_ptr;  const int* h_col;   // CSR over rows: H_H cells (landmark pairs removed)
  const long* off_lloff; const long* off_lnonl; const long* off_other;   // [n1] row offsets
  double* out_lloff; double* out_lnonl; double* out_other;
};

// cells of `cols[lo..hi)` (ascending) that are < j
__device__ __forceinline__ int count_below(const int* __restrict__ cols, int lo, int hi, int j) {
  int c = 0;
  for (int q = lo; q < hi && cols[q] < j; ++q) ++c;
  return c;
}

__global__ void __launch_bounds__(256) separate_scores_kernel(const SepArgs a) {
  const int i = blockIdx.x;
  const bool rl = a.rowL[i] != 0;
  const double* __restrict__ row = a.S + (long)i * a.lds;
  const int plo = a.p_ptr[i], phi = a.p_ptr[i + 1];
  const int hlo = a.h_ptr[i], hhi = a.h_ptr[i + 1];
  double* __restrict__ o_lloff = a.out_lloff + a.off_lloff[i];
  double* __restrict__ o_lnonl = a.out_lnonl + a.off_lnonl[i];
  double* __restrict__ o_other = a.out_other + a.off_other[i];
  // Fast path — every non-landmark row (all but k of the n1 rows), four columns per thread
  // iteration: two 16-byte loads of S, one 4-byte load of the four column flags and one of the
  // rank of the first column.  A chunk without a landmark column and without a homolog cell
  // (~97 % of them) is one contiguous 32-byte piece of "other", shifted by the rank and by the
  // homolog cells before it; the rest go element by element.
  if (!rl && (a.lds & 1) == 0 && (a.n2 & 3) == 0 && ((uintptr_t)a.S & 15) == 0 &&
      ((uintptr_t)a.colL & 3) == 0) {
    const double2* __restrict__ row2 = reinterpret_cast<const double2*>(row);
    const unsigned* __restrict__ cl4 = reinterpret_cast<const unsigned*>(a.colL);
    for (int c = blockIdx.y * blockDim.x + threadIdx.x; c < (a.n2 >> 2); c += gridDim.y * blockDim.x) {
      const double2 v0 = row2[2 * c], v1 = row2[2 * c + 1];
      const unsigned fl = cl4[c];
      const int j = 4 * c;
      const int rk = a.colrank[j];
      int hb = 0;                    // homolog cells of this row in non-landmark columns < j
      int hq = hlo;                  // first homolog cell with column >= j
      for (; hq < hhi; ++hq) {
        const int hc = a.h_col[hq];
        if (hc >= j) break;
        if (!a.colL[hc]) ++hb;
      }
      const bool hin = hq < hhi && a.h_col[hq] < j + 4;
      if (fl == 0u && !hin) {
        double* d = o_other + (j - rk - hb);
        if ((reinterpret_cast<uintptr_t>(d) & 15) == 0) {
          reinterpret_cast<double2*>(d)[0] = v0;
          reinterpret_cast<double2*>(d)[1] = v1;
        } else {
          d[0] = v0.x;
          *reinterpret_cast<double2*>(d + 1) = make_double2(v0.y, v1.x);
          d[3] = v1.y;
        }
      } else {
        const double vv[4] = {v0.x, v0.y, v1.x, v1.y};
        int r = rk;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = j + e;
          const bool is_h = hq < hhi && a.h_col[hq] == col;
          if (is_h) ++hq;
          if ((fl >> (8 * e)) & 0xffu) {
            o_lnonl[r] = vv[e];          // other row, landmark column (homolog cell or not)
            ++r;
          } else if (is_h) {
            ++hb;                        // a homolog cell: not in "other", shifts what follows
          } else {
            o_other[col - r - hb] = vv[e];
          }
        }
      }
    }
    return;
  }
  for (int j = blockIdx.y * blockDim.x + threadIdx.x; j < a.n2; j += gridDim.y * blockDim.x) {
    const double v = row[j];
    const bool cl = a.colL[j] != 0;
    const int rk = a.colrank[j];          // landmark columns before j
    if (rl) {
      if (cl) {
        // landmark row x landmark column: off-diagonal unless (i, j) is a landmark pair
        bool is_pair = false;
        int before = 0;
        for (int q = plo; q < phi; ++q) {
          const int c = a.p_col[q];
          if (c < j) ++before; else { is_pair = (c == j); break; }
        }
        if (!is_pair) o_lloff[rk - before] = v;
      } else {
        o_lnonl[j - rk] = v;               // landmark row, other column
      }
    } else {
      if (cl) {
        o_lnonl[rk] = v;                   // other row, landmark column
      } else {
        // "other" unless the cell is a homolog pair; H_H cells in landmark columns do not shift it
        bool is_h = false;
        int before = 0;
        for (int q = hlo; q < hhi; ++q) {
          const int c = a.h_col[q];
          if (c < j) { if (!a.colL[c]) ++before; } else { is_h = (c == j); break; }
        }
        if (!is_h) o_other[(j - rk) - before] = v;
      }
    }
  }
}

// ---- means only: never materialises the selections -----------------------------------------
// rowsum[i] = sum over non-landmark columns of S[i][:] for non-landmark rows (0 otherwise)
__global__ void __launch_bounds__(256)
masked_row_sum_kernel(int n1, int n2, const double* __restrict__ S, long lds,
                      const unsigned char* __restrict__ rowL, const unsigned char* __restrict__ colL,
                      double* __restrict__ rowsum) {
  const int i = blockIdx.x;
  __shared__ double red[8];
  double s = 0.0;
  if (!rowL[i]) {
    const double* __restrict__ row = S + (long)i * lds;
    for (int j = threadIdx.x; j < n2; j += 256)
      if (!colL[j]) s += row[j];
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    rowsum[i] = t;
  }
}

// out[0] = mean(H_H), out[1] = mean(other), out[2] = out[0] - out[1], out[3] = |H_H|, out[4] = |other|
__global__ void __launch_bounds__(1024)
finish_means_kernel(int n1, const double* __restrict__ rowsum, const double* __restrict__ S, long lds,
                    int nh, const int* __restrict__ hr, const int* __restrict__ hc,
                    const unsigned char* __restrict__ rowL, const unsigned char* __restrict__ colL,
                    double n_other, double* __restrict__ out) {
  __shared__ double red[2][32];
  double tot = 0.0, hh = 0.0, hh_in_other = 0.0;
  for (int i = threadIdx.x; i < n1; i += 1024) tot += rowsum[i];
  for (int q = threadIdx.x; q < nh; q += 1024) {
    const double v = S[(long)hr[q] * lds + hc[q]];
    hh += v;
    if (!rowL[hr[q]] && !colL[hc[q]]) hh_in_other += v;
  }
  tot -= hh_in_other;
  for (int o = 16; o > 0; o >>= 1) {
    tot += __shfl_xor_sync(0xffffffffu, tot, o);
    hh += __shfl_xor_sync(0xffffffffu, hh, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = tot; red[1][threadIdx.x >> 5] = hh; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0, h = 0.0;
    for (int w = 0; w < 32; ++w) { t += red[0][w]; h += red[1][w]; }
    const double hm = h / (double)nh, om = t / n_other;
    out[0] = hm; out[1] = om; out[2] = hm - om; out[3] = (double)nh; out[4] = n_other;
  }
}

}  // namespace
}  // namespace munk

using namespace munk;

extern "C" int munk_separate_scores(int n1, int n2, const double* S, long lds,
                                    const unsigned char* rowL, const unsigned char* colL,
                                    const int* colrank, const int* p_ptr, const int* p_col,
                                    const int* h_ptr, const int* h_col, const long* off_lloff,
                                    const long* off_lnonl, const long* off_other, double* out_lloff,
                                    double* out_lnonl, double* out_other, void* stream) {
  if (n1 <= 0 || n2 <= 0) return MUNK_OK;
  SepArgs a{n1, n2, S, lds, rowL, colL, colrank, p_ptr, p_col, h_ptr, h_col,
            off_lloff, off_lnonl, off_other, out_lloff, out_lnonl, out_other};
  // two blocks per row: each thread streams >= 8 four-column chunks at C3 size, which amortises
  // the per-row set-up (flags, CSR pointers, output offsets) the old 8-way split paid per block
  dim3 grid(n1, (unsigned)std::max(1, std::min(2, (n2 + 2047) / 2048)));
  separate_scores_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  MUNK_CUDA_TRY(cudaGetLastError());
  return MUNK_OK;
}

extern "C" int munk_separate_scores_means(int n1, int n2, const double* S, long lds,
                                          const unsigned char* rowL, const unsigned char* colL,
                                          int nh, const int* h_row, const int* h_col,
                                          double n_other, double* rowsum_scratch, double* out5,
                                          void* stream) {
  if (n1 <= 0 || n2 <= 0) return MUNK_OK;
  masked_row_sum_kernel<<<n1, 256, 0, (cudaStream_t)stream>>>(n1, n2, S, lds, rowL, colL, rowsum_scratch);
  MUNK_CUDA_TRY(cudaGetLastError());
  finish_means_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(n1, rowsum_scratch, S, lds, nh, h_row, h_col,
                                                            rowL, colL, n_other, out5);
  MUNK_CUDA_TRY(cudaGetLastError());
  return MUNK_OK;
}

namespace munk {
namespace {
// out[p] = X[r[p], 0:k] . Z[0:k, c[p]]   — one warp per pair (the score S[r,c] of the factored
// form S = X Z, without forming S).
__global__ void __launch_bounds__(256)
pair_dots_kernel(int npairs, const int* __restrict__ r, const int* __restrict__ c, int k,
                 const double* __restrict__ X, long ldx, const double* __restrict__ Z, long ldz,
                 double* __restrict__ out) {
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= npairs) return;
  const int lane = threadIdx.x & 31;
  const double* __restrict__ x = X + (long)r[p] * ldx;
  const double* __restrict__ z = Z + c[p];
  double s = 0.0;
  for (int l = lane; l < k; l += 32) s += x[l] * z[(long)l * ldz];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[p] = s;
}
}  // namespace
}  // namespace munk

extern "C" int munk_pair_dots(int npairs, const int* r, const int* c, int k, const double* X, long ldx,
                              const double* Z, long ldz, double* out, void* stream) {
  if (npairs <= 0) return MUNK_OK;
  pair_dots_kernel<<<(npairs + 7) / 8, 256, 0, (cudaStream_t)stream>>>(npairs, r, c, k, X, ldx, Z, ldz, out);
  MUNK_CUDA_TRY(cudaGetLastError());
  return MUNK_OK;
}

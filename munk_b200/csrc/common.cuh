// Shared device/host helpers for the munk_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define MUNK_LEAF 128          // diagonal-block size of the recursive factorisations
#define MUNK_MAX_DEVICES 64    // per-device "kernel attribute set" flags
#define MUNK_OK 0

// Status codes of the C ABI (see include/munk_b200.h).
#define MUNK_ERR_CUDA 1000      // + cudaError_t
#define MUNK_ERR_NOT_SPD 2000   // potrf found a non-positive pivot
#define MUNK_ERR_TMAP 3000      // cuTensorMapEncodeTiled failed (+ CUresult)

namespace munk {

void set_last_error(const char* fmt, ...);

#define MUNK_CUDA_TRY(expr)                                                          \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      munk::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                           __FILE__, __LINE__);                                      \
      return MUNK_ERR_CUDA + (int)_e;                                                \
    }                                                                                \
  } while (0)

#define MUNK_TRY(expr)              \
  do {                              \
    int _s = (expr);                \
    if (_s != MUNK_OK) return _s;   \
  } while (0)

static inline long round_up(long x, long m) { return (x + m - 1) / m * m; }

// Makes `device` current for the lifetime of the object (C-ABI entry points: a workspace belongs
// to one device, the caller's current device may be another).
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != device) switched = cudaSetDevice(device) == cudaSuccess;
  }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// ---------------------------------------------------------------------------
// Device-side PTX wrappers: mbarrier, TMA, DMMA.
// ---------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// 2-D tiled TMA load: box at (c0 = innermost coordinate, c1) -> shared, completes on mbarrier.
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(tm), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

// FP64 tensor-core MMA, D(8x8) += A(8x4) * B(4x8).
// lane l holds A[l/4][l%4], B[l%4][l/4], C[l/4][2*(l%4) + {0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
#endif  // __CUDACC__

}  // namespace munk

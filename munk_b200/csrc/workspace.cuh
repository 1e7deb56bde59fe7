// Per-stream scratch owned by the library (opaque `munk_workspace_t` of the C ABI).
#pragma once
#include "common.cuh"

#include <vector>

namespace munk {

struct Workspace {
  int device = 0;
  // split-K partial tiles: [0] for work on the caller's stream, [1] for the look-ahead stream
  double* splitk[2] = {nullptr, nullptr};
  size_t splitk_bytes[2] = {0, 0};
  double* tmp = nullptr;     size_t tmp_bytes = 0;      // trtri off-diagonal temporary
  int* info = nullptr;                                   // device int[8]: potrf status
  int* info_host = nullptr;                              // pinned mirror
  cudaStream_t side = nullptr;                           // high-priority look-ahead stream
  std::vector<cudaEvent_t> events;                       // fork/join + per-panel dependencies
  long launches = 0;                                     // kernels enqueued through this workspace
  double flops = 0.0;                                    // useful flops enqueued

  int ensure(double** p, size_t* cur, size_t need) {
    if (*cur >= need) return MUNK_OK;
    if (*p) MUNK_CUDA_TRY(cudaFree(*p));
    *p = nullptr; *cur = 0;
    MUNK_CUDA_TRY(cudaMalloc((void**)p, need));
    *cur = need;
    return MUNK_OK;
  }
  int ensure_splitk(int which, size_t need) { return ensure(&splitk[which], &splitk_bytes[which], need); }

  // Side stream + `count` events (timing disabled), created once and reused.
  int ensure_side(size_t count) {
    if (!side) {
      int least = 0, greatest = 0;
      MUNK_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&least, &greatest));
      MUNK_CUDA_TRY(cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, greatest));
    }
    while (events.size() < count) {
      cudaEvent_t e;
      MUNK_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      events.push_back(e);
    }
    return MUNK_OK;
  }

  void release() {
    for (int i = 0; i < 2; ++i) if (splitk[i]) cudaFree(splitk[i]);
    if (tmp) cudaFree(tmp);
    if (info) cudaFree(info);
    if (info_host) cudaFreeHost(info_host);
    for (cudaEvent_t e : events) cudaEventDestroy(e);
    if (side) cudaStreamDestroy(side);
  }
};

// Upper bound of the split-K scratch any gemm_launch can ask for (active tiles <= 111,
// active * splits <= 148 + 111, and a lower-only product allocates the full M x N slice, i.e.
// up to twice its active tiles): reserved up front so concurrent streams never reallocate.
constexpr size_t SPLITK_RESERVE = (size_t)530 * 128 * 128 * 8;

}  // namespace munk

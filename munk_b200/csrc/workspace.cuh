// Per-stream scratch owned by the library (opaque `munk_workspace_t` of the C ABI).
#pragma once
#include "common.cuh"

namespace munk {

struct Workspace {
  int device = 0;
  double* splitk = nullptr;  size_t splitk_bytes = 0;   // split-K partial tiles
  double* linv = nullptr;    size_t linv_bytes = 0;     // inverted 128x128 diagonal blocks
  double* tmp = nullptr;     size_t tmp_bytes = 0;      // trtri off-diagonal temporary
  int* info = nullptr;                                   // device int[8]: potrf status
  int* info_host = nullptr;                              // pinned mirror
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
};

}  // namespace munk

// Per-stream scratch owned by the library (opaque `munk_workspace_t` of the C ABI).
#pragma once
#include "common.cuh"

#include <vector>

namespace munk {

constexpr int POOL_STREAMS = 7;                  // extra streams of the forked trtri recursion
constexpr int NSCRATCH = 3 + POOL_STREAMS;       // split-K buffers: caller, side, side2, pool[i]

struct Workspace {
  int device = 0;
  // split-K partial tiles, one buffer per stream that may run GEMMs concurrently:
  // [0] the caller's stream, [1] the diagonal-chain stream, [2] the panel stream, [3+i] pool[i]
  double* splitk[NSCRATCH] = {};
  size_t splitk_bytes[NSCRATCH] = {};
  double* tmp = nullptr;     size_t tmp_bytes = 0;      // trtri off-diagonal temporary
  int* info = nullptr;                                   // device int[8]: [0] potrf status, [1] bad edge
  int* info_host = nullptr;                              // pinned mirror
  cudaStream_t side = nullptr;                           // highest priority: the diagonal-block chain
  cudaStream_t side2 = nullptr;                          // next priority: panel solves + look-ahead updates
  cudaStream_t pool[POOL_STREAMS] = {};                  // independent sub-inversions of trtri
  std::vector<cudaEvent_t> events;                       // fork/join + per-panel dependencies
  // CUDA-graph cache of whole factorisations (factor.cu: graph_or_direct).  A key that was seen
  // once is captured on its second use and replayed afterwards: the ~3 400 launches and ~900
  // event operations of a 20 000-column Cholesky + inverse become two graph launches, so the
  // host is not throttled by the launch queue (measured: 43 ms of blocked host per potrf) and can
  // enqueue the independent target chain at once.
  struct GraphEntry {
    int op, n; const void* a; long ld; const void* linv;
    int n2; const void* a2; long ld2; const void* linv2;      // second matrix of a paired factorisation
    cudaGraphExec_t exec;       // nullptr: seen once, not captured yet
    long launches; double flops;
    unsigned long stamp;
  };
  std::vector<GraphEntry> graphs;
  unsigned long graph_clock = 0;
  long launches = 0;                                     // kernels enqueued through this workspace
  double flops = 0.0;                                    // useful flops enqueued

  int ensure(double** p, size_t* cur, size_t need) {
    if (*cur >= need) return MUNK_OK;
    // a buffer is about to move: captured graphs hold its old address
    if (*p) drop_graphs();
    if (*p) MUNK_CUDA_TRY(cudaFree(*p));
    *p = nullptr; *cur = 0;
    MUNK_CUDA_TRY(cudaMalloc((void**)p, need));
    *cur = need;
    return MUNK_OK;
  }
  void drop_graphs() {
    for (GraphEntry& g : graphs)
      if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
  }
  int ensure_splitk(int which, size_t need) { return ensure(&splitk[which], &splitk_bytes[which], need); }

  // Side stream + `count` events (timing disabled), created once and reused.
  int ensure_side(size_t count) {
    if (!side) {
      int least = 0, greatest = 0;
      MUNK_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&least, &greatest));
      MUNK_CUDA_TRY(cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, greatest));
      const int second = greatest < least ? greatest + 1 : greatest;
      MUNK_CUDA_TRY(cudaStreamCreateWithPriority(&side2, cudaStreamNonBlocking, second));
    }
    return ensure_events(count);
  }

  // Pool streams take the priority of the stream the work was submitted on.
  int ensure_pool(cudaStream_t like) {
    if (pool[0]) return MUNK_OK;
    int prio = 0;
    if (cudaStreamGetPriority(like, &prio) != cudaSuccess) { prio = 0; (void)cudaGetLastError(); }
    for (int i = 0; i < POOL_STREAMS; ++i)
      MUNK_CUDA_TRY(cudaStreamCreateWithPriority(&pool[i], cudaStreamNonBlocking, prio));
    return MUNK_OK;
  }

  int ensure_events(size_t count) {
    while (events.size() < count) {
      cudaEvent_t e;
      MUNK_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      events.push_back(e);
    }
    return MUNK_OK;
  }

  void release() {
    for (int i = 0; i < NSCRATCH; ++i) if (splitk[i]) cudaFree(splitk[i]);
    for (int i = 0; i < POOL_STREAMS; ++i) if (pool[i]) cudaStreamDestroy(pool[i]);
    if (side2) cudaStreamDestroy(side2);
    if (tmp) cudaFree(tmp);
    if (info) cudaFree(info);
    if (info_host) cudaFreeHost(info_host);
    for (cudaEvent_t e : events) cudaEventDestroy(e);
    for (GraphEntry& g : graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    if (side) cudaStreamDestroy(side);
  }
};

// Upper bound of the split-K scratch any gemm_launch can ask for (active tiles <= 111,
// active * splits <= 148 + 111, and a lower-only product allocates the full M x N slice, i.e.
// up to twice its active tiles): reserved up front so concurrent streams never reallocate.
constexpr size_t SPLITK_RESERVE = (size_t)530 * 128 * 128 * 8;

}  // namespace munk

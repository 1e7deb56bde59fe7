// Device -> pageable host memory at memory-bus speed: the output path of the reference-signature
// API (munk.embed_networks / munk.scores return fresh numpy arrays the caller owns, munk/munk/
// munk.py:166-176, scripts/compute_embeddings.py:81).
//
// A fresh pageable array cannot be the target of a fast DMA (cudaMemcpy stages it through one
// driver buffer, single-threaded, and first touch of 8 GB costs seconds), and pinning it in place
// costs more than the copy (2.5 s per 3.2 GB measured on the B200 hosts).  So: T worker threads,
// each with its own pinned staging buffer and CUDA stream, take row chunks round-robin:
//     DMA chunk -> pinned buffer (55 GB/s, the copy engines serialise the chunks)
//     memcpy pinned -> destination (first touch and copy, T threads in parallel)
// While one thread copies out of its buffer the DMA engine fills another thread's.  The pinned
// buffers are allocated once per process and reused.
#include "../../include/munk_b200.h"

#include "common.cuh"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

using namespace munk;

namespace {

constexpr size_t CHUNK_BYTES = (size_t)32 << 20;   // per staging buffer
constexpr int MAX_THREADS = 16;

struct Lane {
  double* pinned = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ready = nullptr;
};

struct Staging {
  std::mutex mu;                 // one download at a time per process (the lanes are shared)
  int device = -1;
  std::vector<Lane> lanes;
  int ensure(int dev, int threads) {
    if (device != dev) {
      release();
      device = dev;
    }
    while ((int)lanes.size() < threads) {
      Lane l;
      MUNK_CUDA_TRY(cudaHostAlloc((void**)&l.pinned, CHUNK_BYTES, cudaHostAllocDefault));
      MUNK_CUDA_TRY(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
      MUNK_CUDA_TRY(cudaEventCreateWithFlags(&l.ready, cudaEventDisableTiming));
      lanes.push_back(l);
    }
    return MUNK_OK;
  }
  void release() {
    for (Lane& l : lanes) {
      if (l.pinned) cudaFreeHost(l.pinned);
      if (l.stream) cudaStreamDestroy(l.stream);
      if (l.ready) cudaEventDestroy(l.ready);
    }
    lanes.clear();
  }
};

Staging& staging() {
  static Staging s;
  return s;
}

}  // namespace

extern "C" int munk_download_pageable(int device, long rows, long cols, const double* src, long lds,
                                      double* host, long ldh, int upper, int threads, void* stream_) {
  if (rows < 0) return -2;
  if (cols < 0) return -3;
  if (rows == 0 || cols == 0) return MUNK_OK;
  if (!src) return -4;
  if (lds < cols) return -5;
  if (!host) return -6;
  if (ldh < cols) return -7;
  if (upper && rows != cols) return -8;
  threads = std::max(1, std::min(threads, MAX_THREADS));
  DeviceGuard guard(device);
  Staging& st = staging();
  std::lock_guard<std::mutex> lock(st.mu);
  MUNK_TRY(st.ensure(device, threads));
  // everything enqueued on `stream` so far must be finished before a lane reads `src`
  cudaEvent_t produced;
  MUNK_CUDA_TRY(cudaEventCreateWithFlags(&produced, cudaEventDisableTiming));
  MUNK_CUDA_TRY(cudaEventRecord(produced, (cudaStream_t)stream_));
  const long chunk_rows = std::max<long>(1, (long)(CHUNK_BYTES / ((size_t)cols * sizeof(double))));
  const long nchunks = (rows + chunk_rows - 1) / chunk_rows;
  std::atomic<long> next(0);
  std::atomic<int> failed(0);
  auto work = [&](int t) {
    cudaSetDevice(device);
    Lane& l = st.lanes[t];
    if (cudaStreamWaitEvent(l.stream, produced, 0) != cudaSuccess) { failed = 1; return; }
    for (;;) {
      const long c = next.fetch_add(1);
      if (c >= nchunks || failed.load()) return;
      const long r0 = c * chunk_rows, nr = std::min(chunk_rows, rows - r0);
      // upper triangular source: columns left of the chunk's first row are zero, not moved
      const long c0 = upper ? (r0 / 16) * 16 : 0;
      const long w = cols - c0;
      cudaError_t e = cudaMemcpy2DAsync(l.pinned, (size_t)w * sizeof(double), src + r0 * lds + c0,
                                        (size_t)lds * sizeof(double), (size_t)w * sizeof(double), (size_t)nr,
                                        cudaMemcpyDeviceToHost, l.stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(l.stream);
      if (e != cudaSuccess) { failed = 1000 + (int)e; return; }
      for (long r = 0; r < nr; ++r) {
        double* dst = host + (r0 + r) * ldh;
        if (c0) memset(dst, 0, (size_t)c0 * sizeof(double));
        memcpy(dst + c0, l.pinned + r * w, (size_t)w * sizeof(double));
      }
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
  work(0);
  for (std::thread& th : pool) th.join();
  cudaEventDestroy(produced);
  if (failed.load()) {
    set_last_error("munk_download_pageable: a staging copy failed (%d)", failed.load());
    return failed.load() >= 1000 ? MUNK_ERR_CUDA + failed.load() - 1000 : MUNK_ERR_CUDA;
  }
  return MUNK_OK;
}

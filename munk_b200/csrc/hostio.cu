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


// ---- content digest of a host buffer --------------------------------------------------------------
// munk.scores() serves np.dot(source_X, target_X.T) from the factors still on the device when it is
// handed the very arrays embed_networks returned (scripts/compute_embeddings.py:72-81).  "The very
// arrays" has to include their contents: the caller owns them and may have edited them in place.
// This digest is taken when the arrays are returned and again when they come back; it reads every
// byte (T threads, memory-bus speed) instead of sampling.
//   chunk (8 MB) digest: four interleaved polynomial hashes h <- h K + w (mod 2^64) over the 64-bit
//   words, folded with an odd multiplier; buffer digest: the chunk digests folded in order (so the
//   result does not depend on the thread count).  Every multiplier is odd, hence invertible mod
//   2^64: changing any single word by d != 0 changes the digest by d * odd != 0 - a one-element
//   edit is always seen; anything else with probability 1 - 2^-64.  out[1] = plain sum of the words.
namespace {

constexpr size_t DIGEST_CHUNK = (size_t)8 << 20;
constexpr unsigned long long DK = 0x9E3779B97F4A7C15ull, DK1 = 0xC2B2AE3D27D4EB4Full, DK2 = 0x165667B19E3779F9ull;

void digest_chunk(const unsigned char* p, size_t bytes, unsigned long long* poly, unsigned long long* sum) {
  unsigned long long h0 = 1, h1 = 2, h2 = 3, h3 = 4, s = 0;
  const size_t words = bytes / 8;
  size_t i = 0;
  for (; i + 4 <= words; i += 4) {
    unsigned long long w[4];
    memcpy(w, p + 8 * i, 32);
    h0 = h0 * DK + w[0];
    h1 = h1 * DK + w[1];
    h2 = h2 * DK + w[2];
    h3 = h3 * DK + w[3];
    s += w[0] + w[1] + w[2] + w[3];
  }
  for (; i < words; ++i) {
    unsigned long long w;
    memcpy(&w, p + 8 * i, 8);
    h0 = h0 * DK + w;
    s += w;
  }
  if (bytes & 7) {
    unsigned long long w = 0;
    memcpy(&w, p + 8 * words, bytes & 7);
    h1 = h1 * DK + w;
    s += w;
  }
  *poly = ((h0 * DK1 + h1) * DK1 + h2) * DK1 + h3;
  *sum = s;
}

}  // namespace

extern "C" int munk_host_digest(const void* data, size_t bytes, int threads, unsigned long long* out) {
  if (!data && bytes) return -1;
  if (!out) return -4;
  const size_t nchunks = (bytes + DIGEST_CHUNK - 1) / DIGEST_CHUNK;
  threads = (int)std::max<size_t>(1, std::min<size_t>({(size_t)std::max(threads, 1), (size_t)64, std::max<size_t>(nchunks, 1)}));
  std::vector<unsigned long long> poly(nchunks), sum(nchunks);
  std::atomic<size_t> next(0);
  const unsigned char* base = static_cast<const unsigned char*>(data);
  auto work = [&]() {
    for (;;) {
      const size_t c = next.fetch_add(1);
      if (c >= nchunks) return;
      const size_t off = c * DIGEST_CHUNK;
      digest_chunk(base + off, std::min(DIGEST_CHUNK, bytes - off), &poly[c], &sum[c]);
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < threads; ++t) pool.emplace_back(work);
  work();
  for (std::thread& th : pool) th.join();
  unsigned long long h = bytes, s = 0;
  for (size_t c = 0; c < nchunks; ++c) {
    h = h * DK2 + poly[c];
    s += sum[c];
  }
  out[0] = h;
  out[1] = s;
  return MUNK_OK;
}

// Register-only micro-benchmarks: the FP64 ceilings every roofline fraction is quoted against.
// (MEASURED_PEAKS.json has no FP64 entry; SURVEY.md section 6 asks for it to be measured.)
#include "../../include/munk_b200.h"
#include "common.cuh"

namespace munk {
namespace {

// 8 warps x 32 independent DMMA accumulators per warp, no memory traffic.
__global__ void __launch_bounds__(256) dmma_peak_kernel(int iters, double* sink) {
  double acc[32][2];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i][0] = acc[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) dmma884(acc[i][0], acc[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += acc[i][0] + acc[i][1];
  if (s == 123.456) sink[0] = s;
}

__global__ void __launch_bounds__(256) dfma_peak_kernel(int iters, double* sink) {
  double acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = i;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += acc[i];
  if (s == 123.456) sink[0] = s;
}

template <typename K>
int run(K kern, int device, int iters, double flops_per_thread_iter, double* tflops) {
  MUNK_CUDA_TRY(cudaSetDevice(device));
  double* sink = nullptr;
  MUNK_CUDA_TRY(cudaMalloc((void**)&sink, 8));
  cudaEvent_t e0, e1;
  MUNK_CUDA_TRY(cudaEventCreate(&e0));
  MUNK_CUDA_TRY(cudaEventCreate(&e1));
  const int blocks = 148 * 4, threads = 256;
  kern<<<blocks, threads>>>(iters / 8 + 1, sink);   // warm-up
  MUNK_CUDA_TRY(cudaEventRecord(e0));
  kern<<<blocks, threads>>>(iters, sink);
  MUNK_CUDA_TRY(cudaEventRecord(e1));
  MUNK_CUDA_TRY(cudaEventSynchronize(e1));
  float ms = 0.f;
  MUNK_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
  *tflops = flops_per_thread_iter * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  return MUNK_OK;
}

}  // namespace
}  // namespace munk

extern "C" int munk_microbench_dmma(int device, int iters, double* tflops) {
  // one m8n8k4 DMMA = 8*8*4 FMA = 512 flops per warp = 16 per thread; 32 per iteration
  return munk::run(munk::dmma_peak_kernel, device, iters, 32.0 * 16.0, tflops);
}
extern "C" int munk_microbench_dfma(int device, int iters, double* tflops) {
  return munk::run(munk::dfma_peak_kernel, device, iters, 32.0 * 2.0, tflops);
}

// fp32_peak.cu -- measurement helper for bench.py (NOT part of the product library).
//
// MEASURED_PEAKS.json has HBM and bf16 tensor peaks only; the many-chain log-density kernel is
// bound by non-tensor FP32 issue / shared-memory bandwidth, so bench.py measures
// an FFMA peak and a shared-memory read peak on the same GPU, in the same run, as denominators.
#include <cuda_runtime.h>
#include <stdint.h>

extern "C" {

__global__ void __launch_bounds__(1024) ffma_kernel(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll 16
    for (int j = 0; j < 16; j++) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// each warp streams conflict-free 256-byte rows (LDS.64) out of a 64 KB shared buffer
__global__ void __launch_bounds__(1024) lds_kernel(float* out, int iters) {
  extern __shared__ float2 sm2[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm2[i] = make_float2(1.0f, 2.0f);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  unsigned row = threadIdx.x >> 5;
  float sx = 0.f, sy = 0.f;
  for (int i = 0; i < iters; i++) {
#pragma unroll 16
    for (int j = 0; j < 16; j++) {
      const float2 v = sm2[((row + j * 7) & 255) * 32 + lane];
      sx += v.x;
      sy += v.y;
    }
    row += 113;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = sx + sy;
}

// returns best-of-reps TFLOP/s (FMA = 2 flops) of dependent-chain FFMA issue on the current device
double bplxbench_fp32_peak(int iters, int reps) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int blocks = sms * 2, threads = 1024;
  float* out = nullptr;
  if (cudaMalloc(&out, (size_t)blocks * threads * sizeof(float)) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int r = 0; r < reps + 2; r++) {
    cudaEventRecord(e0);
    ffma_kernel<<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 8.0 * 16.0 * (double)iters * blocks * threads;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (r >= 2 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return cudaGetLastError() == cudaSuccess ? best : -1.0;
}

// returns best-of-reps shared-memory read bandwidth in TB/s (all SMs)
double bplxbench_smem_peak(int iters, int reps) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int blocks = sms * 2, threads = 1024;
  float* out = nullptr;
  if (cudaMalloc(&out, (size_t)blocks * threads * sizeof(float)) != cudaSuccess) return -1.0;
  cudaFuncSetAttribute(lds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int r = 0; r < reps + 2; r++) {
    cudaEventRecord(e0);
    lds_kernel<<<blocks, threads, 65536>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = 8.0 * 16.0 * (double)iters * blocks * threads;
    const double tb = bytes / (ms * 1e-3) / 1e12;
    if (r >= 2 && tb > best) best = tb;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return cudaGetLastError() == cudaSuccess ? best : -1.0;
}

}  // extern "C"

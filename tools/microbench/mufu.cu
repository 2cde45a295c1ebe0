// Microbenchmark: MUFU.EX2 and FFMA issue rates per SM on this GPU.  nvcc -arch=sm_100a -O3 mufu.cu -o mufu && ./mufu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_ex2(float* out, int iters) {
  float v[8];
  for (int i = 0; i < 8; ++i) v[i] = -0.001f * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
  float s = 0; for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_fma(float* out, int iters) {
  float v[8];
  for (int i = 0; i < 8; ++i) v[i] = -0.001f * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(v[i]));
  float s = 0; for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  int dev = 0, sms = 0, khz = 0; cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  float* out; cudaMalloc(&out, sizeof(float) * sms * 1024);
  for (int warps = 4; warps <= 32; warps *= 2) {
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int which = 0; which < 2; ++which) {
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (which == 0) k_ex2<<<sms, warps * 32>>>(out, iters); else k_fma<<<sms, warps * 32>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double ops = (double)warps * 32 * 8 * iters;          // per SM
      double cyc = ms * 1e-3 * khz * 1e3;                   // at max clock
      printf("%s warps/SM=%2d: %.3f ms  -> %.1f ops/clk/SM (assuming %d MHz)\n", which == 0 ? "ex2" : "fma", warps, ms, ops / cyc, khz / 1000);
    }
  }
  return 0;
}

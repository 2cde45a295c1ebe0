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
// packed half: two exponentials per lane and instruction (if the SFU really delivers both per issue slot)
__global__ void k_ex2h2(float* out, int iters) {
  unsigned v[8];
  for (int i = 0; i < 8; ++i) v[i] = 0xB800B800u + threadIdx.x + i;   // two small negative halves
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(v[i]));
  unsigned s = 0; for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(s);
}
__global__ void k_ex2b2(float* out, int iters) {
  unsigned v[8];
  for (int i = 0; i < 8; ++i) v[i] = 0xBF00BF00u + threadIdx.x + i;   // two small negative bf16
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(v[i]));
  unsigned s = 0; for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(s);
}
// fp32 pair -> packed bf16x2 (F2FP.BF16.PACK_AB): which pipe, what rate?
__global__ void k_f2fp(float* out, int iters) {
  float v[8];
  unsigned acc = 0;
  for (int i = 0; i < 8; ++i) v[i] = 1.0f + 0.001f * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      unsigned r;
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(v[i]), "f"(v[(i + 1) & 7]));
      acc ^= r;                                     // one LOP per conversion keeps the result alive
    }
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
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
    for (int which = 0; which < 5; ++which) {
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (which == 0) k_ex2<<<sms, warps * 32>>>(out, iters); else if (which == 1) k_fma<<<sms, warps * 32>>>(out, iters); else if (which == 2) k_ex2h2<<<sms, warps * 32>>>(out, iters); else if (which == 3) k_ex2b2<<<sms, warps * 32>>>(out, iters); else k_f2fp<<<sms, warps * 32>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double ops = (double)warps * 32 * 8 * iters;          // per SM
      double cyc = ms * 1e-3 * khz * 1e3;                   // at max clock
      printf("%s warps/SM=%2d: %.3f ms  -> %.1f ops/clk/SM (assuming %d MHz)\n", which == 0 ? "ex2" : which == 1 ? "fma" : which == 2 ? "ex2.f16x2 (instr)" : which == 3 ? "ex2.bf16x2 (instr)" : "cvt.rn.bf16x2.f32 (instr)", warps, ms, ops / cyc, khz / 1000);
    }
  }
  return 0;
}

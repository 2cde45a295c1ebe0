// Microbenchmark: issue rates of the instructions of the attention softmax loop, alone and mixed, in cycles measured on the
// SM itself (clock64), for 1, 2 and 4 warps per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 pipes.cu -o pipes && ./pipes
// Prints warp-instructions per clock per sub-partition (1.0 = the scheduler's limit) and, for the mixes, cycles per "softmax
// element pair" so that the loop's pipe budget can be read off directly.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define N 8
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }

template <int KIND>
__global__ void k(float* out, long long* cyc, int iters) {
  float v[N], w[N];
  u64 p[N], q[N];
  unsigned acc = 0;
  for (int i = 0; i < N; ++i) { v[i] = -0.001f * (threadIdx.x + i + 1); w[i] = 1.0f + 0.01f * i; p[i] = pk(v[i], w[i]); q[i] = pk(w[i], v[i]); }
  const u64 c1 = pk(0.999f, 1.001f), c2 = pk(-0.5f, 0.25f);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (KIND == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(w[i]), "f"(w[(i + 1) % N]));                 // FFMA, 3 registers
      if (KIND == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(c1), "l"(c2));                              // FFMA2
      if (KIND == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(q[i]));                                         // FADD2
      if (KIND == 3) asm volatile("max.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(w[i]));                                              // FMNMX
      if (KIND == 4) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(w[i]), "f"(w[(i + 1) % N]));                     // FMNMX3
      if (KIND == 5) { unsigned r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(v[i]), "f"(w[i])); acc ^= r; }  // F2FP (+LOP)
      if (KIND == 6) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));                                                   // MUFU.EX2
      if (KIND == 7) {   // the loop's mix per element PAIR on the SFU path: FFMA2 (scale), 2 x EX2, FADD2 (sum), F2FP (pack), 1/2 FMNMX3 (max)
        float a, b;
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(c1), "l"(c2));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p[i]));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(q[i]) : "l"(pk(a, b)));
        unsigned r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b)); acc ^= r;
        if (i & 1) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(a), "f"(b));
      }
      if (KIND == 8) {   // the same mix with scalar fp32 instead of the packed instructions
        float a, b;
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p[i]));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a) : "f"(w[i]), "f"(w[(i + 1) % N]));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(b) : "f"(w[i]), "f"(w[(i + 1) % N]));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b));
        asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(a));
        asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(v[(i + 1) % N]) : "f"(b));
        unsigned r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b)); acc ^= r;
        if (i & 1) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(w[i]) : "f"(a), "f"(b));
      }
    }
  }
  const long long t1 = clock64();
  float s = __uint_as_float(acc);
  for (int i = 0; i < N; ++i) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p[i])); s += v[i] + w[i] + a + b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(q[i])); s += a + b; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int KIND>
static void run(const char* name, int per_iter_instr, float* out, long long* cyc, int sms) {
  for (int wps = 1; wps <= 4; wps *= 2) {          // warps per sub-partition
    const int iters = 4000;
    k<KIND><<<sms, wps * 4 * 32>>>(out, cyc, iters);
    k<KIND><<<sms, wps * 4 * 32>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long c[8]; cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    const double cycles = (double)c[0];
    const double winstr = (double)wps * N * iters * per_iter_instr;    // per sub-partition
    printf("%-34s %d warp/SMSP: %.3f warp-instr/clk/SMSP, %.2f clk per unrolled item per warp\n", name, wps, winstr / cycles, cycles / (N * (double)iters));
  }
}

int main() {
  int dev = 0, sms = 0; cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  float* out; cudaMalloc(&out, sizeof(float) * sms * 1024);
  long long* cyc; cudaMalloc(&cyc, sizeof(long long) * sms);
  run<0>("FFMA (3 regs)", 1, out, cyc, sms);
  run<1>("FFMA2 (fma.rn.f32x2)", 1, out, cyc, sms);
  run<2>("FADD2 (add.rn.f32x2)", 1, out, cyc, sms);
  run<3>("FMNMX", 1, out, cyc, sms);
  run<4>("FMNMX3", 1, out, cyc, sms);
  run<5>("F2FP.BF16 (+LOP)", 2, out, cyc, sms);
  run<6>("MUFU.EX2", 1, out, cyc, sms);
  run<7>("softmax mix, packed (per pair)", 6, out, cyc, sms);
  run<8>("softmax mix, scalar (per pair)", 8, out, cyc, sms);
  return 0;
}

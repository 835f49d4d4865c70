// Issue-rate microbenchmark of the instruction mixes the field kernels are made of (sm_100a).
// Each kernel runs ITER iterations of UNROLL independent dependency chains per thread; reports
// warp-instructions per clock per SM sub-partition.   nvcc -O3 -arch=sm_100a -o pipes pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
#define ITER 4096
#define CH 8

template <int MODE> __global__ void k(u64 *out, u32 seed) {
  u64 a[CH]; u32 b[CH]; double d[CH];
  for (int i = 0; i < CH; i++) { a[i] = seed * (i + 3) + threadIdx.x; b[i] = seed + i * 7 + threadIdx.x; d[i] = (double)(seed + i); }
  u32 m = seed | 1; double dm = (double)(seed & 0xffff);
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) {
      if (MODE == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[i]) : "r"(b[i]), "r"(m));          // IMAD.WIDE.U32
      if (MODE == 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(m), "r"(seed));            // IMAD
      if (MODE == 2) asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(m));                               // IADD3
      if (MODE == 3) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(dm), "d"(d[(i + 1) % CH])); // DFMA
      if (MODE == 4) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[i]) : "r"(b[i]), "r"(m));       // IMAD.WIDE + IADD3 pair
                       asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(m)); }
      if (MODE == 5) { asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(dm), "d"(d[(i + 1) % CH])); // DFMA + IADD3 + IMAD.WIDE
                       asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(m));
                       asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[i]) : "r"(b[i]), "r"(m)); }
      if (MODE == 6) { u32 lo = (u32)a[i], hi = (u32)(a[i] >> 32);                                           // 64-bit add with carry (IADD3 + IADD3.X)
                       asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(lo), "+r"(hi) : "r"(m), "r"(seed));
                       a[i] = ((u64)hi << 32) | lo; }
      if (MODE == 7) b[i] = __shfl_sync(0xffffffffu, b[i], (threadIdx.x + 1) & 31);                           // SHFL
      if (MODE == 8) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(m), "r"(seed));         // LOP3
      if (MODE == 9) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(m), "r"(seed));          // IMAD + IADD3
                       asm volatile("add.u32 %0, %0, %1;" : "+r"(b[(i + 1) % CH]) : "r"(m)); }
    }
  }
  u64 r = 0;
  for (int i = 0; i < CH; i++) r += a[i] + b[i] + (u64)d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE> void run(const char *name, int per_iter) {
  int sms = 148, threads = 512, blocks = sms * 4;
  u64 *out; cudaMalloc(&out, (size_t)blocks * threads * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(out, 12345);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(out, 12345);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double warp_instr = (double)blocks * threads / 32 * ITER * CH * per_iter;
  double clocks = ms * 1e-3 * 1.965e9;
  printf("%-34s %8.3f ms  %6.3f warp-instr/clk/SMSP (assuming 1965 MHz)\n", name, ms, warp_instr / clocks / (sms * 4));
  cudaFree(out);
}
int main() {
  run<0>("IMAD.WIDE.U32", 1); run<1>("IMAD (32-bit)", 1); run<2>("IADD3", 1); run<3>("DFMA", 1);
  run<4>("IMAD.WIDE + IADD3", 2); run<5>("DFMA + IADD3 + IMAD.WIDE", 3); run<6>("IADD3 + IADD3.X (64-bit add)", 2);
  run<7>("SHFL.IDX", 1); run<8>("LOP3", 1); run<9>("IMAD + IADD3", 2);
  return 0;
}

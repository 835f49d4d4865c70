// Throughput of the field primitives and of the Tip5 MDS step in isolation (sm_100a), to compare with the
// instruction-count bound (pipes.cu: IMAD/IMAD.WIDE 0.5, IADD3 1.0, IADD3.X/LOP3 0.5 warp-instr/clk/SMSP).
#include <cstdio>
#include "../../triton-vm_b200/csrc/field.cuh"
using namespace tvm;
#define ITER 2048
template <int MODE, int CH> __global__ void __launch_bounds__(128) k(u64 *out, u64 seed) {
  u64 a[CH], b[CH];
  for (int i = 0; i < CH; i++) { a[i] = seed * (i + 3) + threadIdx.x; b[i] = (seed + i * 7 + threadIdx.x) | 1; }
  const int l = threadIdx.x & 3, gb = threadIdx.x & 28;
  for (int it = 0; it < ITER; it++) {
    if (MODE == 0) { for (int i = 0; i < CH; i++) a[i] = fmul(a[i], b[i]); }
    if (MODE == 1) { for (int i = 0; i < CH; i++) a[i] = fadd(a[i], b[i]); }
    if (MODE == 2) { for (int i = 0; i < CH; i++) a[i] = fsub(a[i], b[i]); }
    if (MODE == 3) { for (int i = 0; i < CH; i++) { u64 t = fmul(a[i], b[i]); a[i] = fadd(t, b[i]); b[i] = fsub(t, a[i]); } }  // butterfly-like
    if (MODE == 4) {  // MDS accumulate of the quad kernel: 16 shuffles, 128 IMAD.WIDE
      constexpr unsigned short MDS[16] = {61402, 1108, 28750, 33823, 7454, 43244, 53865, 12034, 56951, 27521, 41351, 40901, 12021, 59689, 26798, 17845};
      u64 lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};
#pragma unroll
      for (int k2 = 0; k2 < 16; k2++) {
        const int wrap = (((l - k2) & 3) + (k2 & 3)) >> 2;
        const u64 provide = wrap ? a[((k2 >> 2) + 1) & 3] : a[k2 >> 2];
        unsigned plo = __shfl_sync(0xffffffffu, (unsigned)provide, gb + ((k2 + l) & 3));
        unsigned phi = __shfl_sync(0xffffffffu, (unsigned)(provide >> 32), gb + ((k2 + l) & 3));
#pragma unroll
        for (int i = 0; i < 4; i++) { const u64 m = MDS[(4 * i - k2) & 15]; lo[i] += m * plo; hi[i] += m * phi; }
      }
      for (int i = 0; i < 4; i++) { u64 ls = lo[i] + (hi[i] << 32); u64 cy = ls < lo[i]; a[i] = reduce96(ls, (hi[i] >> 32) + cy); }
    }
    if (MODE == 5) {  // x^7 on 3 elements (S-box power map of one lane)
      for (int i = 1; i < 4; i++) { u64 x = a[i], x2 = fmul(x, x), x3 = fmul(x2, x), x4 = fmul(x2, x2); a[i] = fmul(x3, x4); }
    }
  }
  u64 r = 0;
  for (int i = 0; i < CH; i++) r += a[i] ^ b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE, int CH> void run(const char *name, double ops_per_iter) {
  int blocks = 148 * 9;
  u64 *out; cudaMalloc(&out, (size_t)blocks * 128 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE, CH><<<blocks, 128>>>(out, 12345);
  cudaEventRecord(e0);
  k<MODE, CH><<<blocks, 128>>>(out, 12345);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double warp_ops = (double)blocks * 4 * ITER * ops_per_iter;
  double clk_per_op = (ms * 1e-3 * 1.965e9) * (148 * 4) / warp_ops;
  printf("%-44s %8.3f ms   %7.2f SMSP-clocks per warp-op\n", name, ms, clk_per_op);
  cudaFree(out);
}
int main() {
  run<0, 8>("fmul (8 independent chains)", 8); run<0, 4>("fmul (4 chains)", 4); run<1, 8>("fadd", 8); run<2, 8>("fsub", 8);
  run<3, 4>("fmul+fadd+fsub (butterfly, 4 chains)", 4); run<4, 4>("Tip5 MDS step of one lane (per round)", 1);
  run<5, 4>("x^7 on 3 elements (per round)", 1);
  return 0;
}

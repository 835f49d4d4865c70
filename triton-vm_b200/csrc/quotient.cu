// Quotient codeword: host-side driver of the generated AIR kernels.
// Restates Prover::compute_quotient_segments' cached branch up to the quotient codeword
// (triton-vm/src/stark.rs:770-782) = all_quotients_combined (master_table.rs:1264-1363).
#include "air.cuh"
#include "launch.h"
#include "air_gen/air_meta.inc"

namespace tvm {

#define TVM_AIR_TU(name) void name##_launch(const AirArgs &a, const u64 *d_wtab, const u64 *d_ch, cudaStream_t s, unsigned long long *launches);
#include "air_gen/air_chunks.inc"
#undef TVM_AIR_TU

// weight table: 7 words per weight (air.cuh, AIR_WTAB_WORDS)
__global__ void air_weight_table_kernel(const u64 *w, unsigned count, u64 *tab) {
  unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  u64 b0 = w[3 * j], b1 = w[3 * j + 1], b2 = w[3 * j + 2];
  u64 *t = tab + (size_t)AIR_WTAB_WORDS * j;
  t[0] = b0; t[1] = b1; t[2] = b2; t[3] = fneg(b1); t[4] = fneg(b2); t[5] = fadd(b0, b2); t[6] = fsub(b1, b2);
}

// Per-row zerofier inverses.  Each thread owns ZF_PER rows and inverts
// their (x - 1) and (x - w_n^-1) with one shared field inversion (Montgomery's trick).
static constexpr int ZF_PER = 8;
struct ZerofierArgs {
  u64 *zi_init, *zi_tran, *zi_term;
  size_t nrows;
  int log_n;
  PowTab trace_gen;
  u64 trace_gen_inv;
  u64 coset_x[AIR_MAX_COSETS];
  u64 cons_zerofier_inv[AIR_MAX_COSETS];
};
__global__ void __launch_bounds__(256) air_zerofier_kernel(ZerofierArgs a) {
  // the ZF_PER rows of a thread are strided by the total thread count so that stores stay coalesced
  const size_t nthreads = (size_t)gridDim.x * blockDim.x;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n = (size_t)1 << a.log_n;
  u64 d[2 * ZF_PER], pre[2 * ZF_PER];
  u64 run = MONT_ONE;
#pragma unroll
  for (int j = 0; j < ZF_PER; j++) {
    size_t m = t + (size_t)j * nthreads;
    if (m >= a.nrows) m = a.nrows - 1;
    size_t coset = m >> a.log_n, k = m & (n - 1);
    u64 lo = __ldg(a.trace_gen.lo + (k & ((1ULL << a.trace_gen.shift) - 1)));
    u64 hi = __ldg(a.trace_gen.hi + (k >> a.trace_gen.shift));
    u64 x = fmul(a.coset_x[coset], fmul(lo, hi));
    d[2 * j] = fsub(x, MONT_ONE);
    d[2 * j + 1] = fsub(x, a.trace_gen_inv);
    pre[2 * j] = run; run = fmul(run, d[2 * j]);
    pre[2 * j + 1] = run; run = fmul(run, d[2 * j + 1]);
  }
  u64 inv = finv(run);
#pragma unroll
  for (int j = ZF_PER - 1; j >= 0; j--) {
    u64 i1 = fmul(inv, pre[2 * j + 1]); inv = fmul(inv, d[2 * j + 1]);
    u64 i0 = fmul(inv, pre[2 * j]);     inv = fmul(inv, d[2 * j]);
    size_t m = t + (size_t)j * nthreads;
    if (m < a.nrows) {
      a.zi_init[m] = i0;
      a.zi_term[m] = i1;
      a.zi_tran[m] = fmul(d[2 * j + 1], a.cons_zerofier_inv[m >> a.log_n]);
    }
  }
}

// d_main: [>=379][r*n] coset-major LDE of the main table; d_aux: [>=270][r*n] (X-field columns as
// 3 planar B-field columns); d_challenges: 63 X-field (Montgomery); d_weights: 604 X-field;
// d_out: 3 planes of r*n words (memory order = coset-major), overwritten.
void air_quotient_run(Ctx &c, const u64 *d_main, size_t main_stride, const u64 *d_aux, size_t aux_stride,
                      const u64 *d_challenges, const u64 *d_weights, unsigned log_n, unsigned log_r,
                      unsigned coset_first, unsigned coset_step, unsigned num_cosets, unsigned coset_mem_stride,
                      u64 offset_mont, u64 *d_out, size_t out_stride) {
  if (num_cosets > (unsigned)AIR_MAX_COSETS) throw ApiError{TVM_ERR_UNSUPPORTED, "too many cosets"};
  const unsigned num_weights = AIR_NUM_INIT + AIR_NUM_CONS + AIR_NUM_TRAN + AIR_NUM_TERM;
  AirArgs a{};
  a.main = d_main; a.main_stride = main_stride; a.aux = d_aux; a.aux_stride = aux_stride;
  a.out = d_out; a.out_stride = out_stride;
  a.nrows = (size_t)num_cosets << log_n;
  a.log_n = (int)log_n;
  a.coset_mem_stride = coset_mem_stride;
  ZerofierArgs z{};
  z.nrows = a.nrows; z.log_n = a.log_n;
  u64 wn = root_of_unity_mont(log_n);
  z.trace_gen = c.get_pow_tab(wn, (int)log_n);
  z.trace_gen_inv = finv(wn);
  u64 wrn = root_of_unity_mont(log_n + log_r);
  for (unsigned y = 0; y < num_cosets; y++) {
    u64 x = fmul(offset_mont, fpow(wrn, (u64)(coset_first + coset_step * y)));
    z.coset_x[y] = x;
    z.cons_zerofier_inv[y] = a.cons_zerofier_inv[y] = finv(fsub(fpow(x, (u64)1 << log_n), MONT_ONE));
  }
  u64 *scratch = (u64 *)c.pool_alloc(sizeof(u64) * (3 * a.nrows + (size_t)AIR_WTAB_WORDS * num_weights));
  z.zi_init = scratch; z.zi_tran = scratch + a.nrows; z.zi_term = scratch + 2 * a.nrows;
  u64 *d_wtab = scratch + 3 * a.nrows;
  a.zi_init = z.zi_init; a.zi_tran = z.zi_tran; a.zi_term = z.zi_term;
  try {
    size_t zthreads = (a.nrows + ZF_PER - 1) / ZF_PER;
    air_zerofier_kernel<<<(unsigned)((zthreads + 255) / 256), 256, 0, c.stream>>>(z);
    air_weight_table_kernel<<<(num_weights + 127) / 128, 128, 0, c.stream>>>(d_weights, num_weights, d_wtab);
    c.launches += 2;
    TVM_CUDA(cudaMemsetAsync(d_out, 0, sizeof(u64) * a.nrows, c.stream));
    TVM_CUDA(cudaMemsetAsync(d_out + out_stride, 0, sizeof(u64) * a.nrows, c.stream));
    TVM_CUDA(cudaMemsetAsync(d_out + 2 * out_stride, 0, sizeof(u64) * a.nrows, c.stream));
#define TVM_AIR_TU(name) name##_launch(a, d_wtab, d_challenges, c.stream, &c.launches);
#include "air_gen/air_chunks.inc"
#undef TVM_AIR_TU
    TVM_CUDA(cudaGetLastError());
  } catch (...) {
    c.pool_release(scratch);
    throw;
  }
  c.pool_release(scratch);
}

}  // namespace tvm

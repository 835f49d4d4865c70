// Quotient codeword: host-side driver of the generated AIR kernels.
// Restates Prover::compute_quotient_segments' cached branch up to the quotient codeword
// (triton-vm/src/stark.rs:770-782) = all_quotients_combined (master_table.rs:1264-1363).
#include "air.cuh"
#include "launch.h"
#include "air_gen/air_meta.inc"
#include "stark.h"
#include <cstdlib>

namespace tvm {

#define TVM_AIR_TU(name) void name##_launch(const AirArgs &a, const AirArgs &lo, const u64 *d_wtab, const u64 *d_ch, cudaStream_t s, unsigned long long *launches);
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

// Degree split, last step: out += sum over the four zerofier classes of zerofier_inverse_class(row) * S_class(row), where S_class
// (the weighted sum of the class's constraints of degree <= 2) sits in coset-major order for the even cosets (computed by the
// "lo" chunks) and in NATURAL order of the half-size domain for the odd cosets (extended by two transforms, air_quotient_run).
__global__ void __launch_bounds__(256) air_low_combine_kernel(AirArgs a, const u64 *s_even, const u64 *s_odd, size_t ls, int log_half) {
  const size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= a.nrows) return;
  const size_t n = (size_t)1 << a.log_n;
  const size_t coset = m >> a.log_n, k = m & (n - 1), half = coset >> 1;
  const u64 *src = (coset & 1) ? s_odd + half + (k << log_half) : s_even + half * n + k;
  const u64 zi[4] = {a.zi_init[m], a.cons_zerofier_inv[coset], a.zi_tran[m], a.zi_term[m]};
  xfe acc = xzero();
#pragma unroll
  for (int t = 0; t < 4; t++) acc = xadd(acc, xmulb(xmake(src[(3 * t) * ls], src[(3 * t + 1) * ls], src[(3 * t + 2) * ls]), zi[t]));
  air_add_out(a, m, acc);
}

// d_main: [>=379][r*n] coset-major LDE of the main table; d_aux: [>=270][r*n] (X-field columns as
// 3 planar B-field columns); d_challenges: 63 X-field (Montgomery); d_weights: 604 X-field;
// d_out: 3 planes of r*n words (memory order = coset-major), overwritten.
void air_quotient_run(Ctx &c, const u64 *d_main, size_t main_stride, const u64 *d_aux, size_t aux_stride,
                      const u64 *d_challenges, const u64 *d_weights, unsigned log_n, unsigned log_r,
                      unsigned coset_first, unsigned coset_step, unsigned num_cosets, unsigned coset_mem_stride,
                      u64 offset_mont, u64 *d_out, size_t out_stride, bool low_degree_tables) {
  if (num_cosets > (unsigned)AIR_MAX_COSETS) throw ApiError{TVM_ERR_UNSUPPORTED, "too many cosets"};
  // the degree split needs the whole domain in one call, at least 8n points of it (the class sums have degree < 4n), and tables
  // whose columns are polynomials of degree < 2n (true for the prover's extended tables, not for arbitrary test data)
  static const bool split_enabled = [] { const char *e = getenv("TVM_AIR_NO_DEGREE_SPLIT"); return !(e && e[0] == '1'); }();
  const bool split = TVM_AIR_HAS_LOW_CHUNKS && split_enabled && low_degree_tables && coset_first == 0 && coset_step == 1 &&
                     num_cosets == (1u << log_r) && log_r >= 3 && log_n + log_r - 1 <= 26;
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
  u64 *low = nullptr;
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
    AirArgs lo = a;
    const size_t ls = a.nrows / 2;               // rows of the even cosets = plane stride of the class sums
    if (split) {
      low = (u64 *)c.pool_alloc(sizeof(u64) * 48 * ls);   // 12 planes each: class sums | natural order, later odd values | coefficients | NTT scratch
      TVM_CUDA(cudaMemsetAsync(low, 0, sizeof(u64) * 12 * ls, c.stream));
      lo.nrows = ls;
      lo.coset_mem_stride = 2 * coset_mem_stride;           // evaluated coset y of the lo chunks = domain coset 2y
      lo.low_out = low; lo.low_stride = ls;
    }
#define TVM_AIR_TU(name) name##_launch(a, lo, d_wtab, d_challenges, c.stream, &c.launches);
#include "air_gen/air_chunks.inc"
#undef TVM_AIR_TU
    TVM_CUDA(cudaGetLastError());
    if (split) {
      // even cosets = the coset offset*<w_{rn}^2> of `ls` points: natural order, inverse transform (= coefficients times
      // offset^j), times w_{rn}^j, forward transform = the values on offset*w_{rn}*<w_{rn}^2>, the odd cosets, natural order
      u64 *nat = low + 12 * ls, *coef = low + 24 * ls, *ntt_tmp = low + 36 * ls;
      const int log_half = (int)log_r - 1, log_len = (int)(log_n + log_r) - 1;
      coset_to_natural_run(c, low, nat, ls, ls, (int)log_n, log_half, 12);
      NttJob inv{};
      inv.in = nat; inv.in_cstride = ls; inv.out = coef; inv.out_cstride = ls; inv.tmp = ntt_tmp;
      inv.log_n = log_len; inv.ncols = 12; inv.inverse = true;
      ntt_run(c, inv);
      scale_by_powers_run(c, coef, ls, 12, ls, c.get_pow_tab(wrn, log_len));
      NttJob fwd{};
      fwd.in = coef; fwd.in_cstride = ls; fwd.out = nat; fwd.out_cstride = ls; fwd.tmp = ntt_tmp;
      fwd.log_n = log_len; fwd.ncols = 12; fwd.inverse = false;
      ntt_run(c, fwd);
      air_low_combine_kernel<<<(unsigned)((a.nrows + 255) / 256), 256, 0, c.stream>>>(a, low, nat, ls, log_half);
      c.launches++;
      TVM_CUDA(cudaGetLastError());
    }
  } catch (...) {
    c.pool_release(scratch);
    if (low) c.pool_release(low);
    throw;
  }
  c.pool_release(scratch);
  if (low) c.pool_release(low);
}

}  // namespace tvm

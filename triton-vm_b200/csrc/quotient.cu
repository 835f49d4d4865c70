// Quotient codeword: host-side driver of the generated AIR kernels.
// Restates Prover::compute_quotient_segments' cached branch up to the quotient codeword
// (triton-vm/src/stark.rs:770-782) = all_quotients_combined (master_table.rs:1264-1363).
#include "air.cuh"
#include "launch.h"

namespace tvm {

#define TVM_AIR_CHUNK(name) void name##_launch(const AirArgs &a, const u64 *d_w, const u64 *d_ch, cudaStream_t s);
#include "air_gen/air_chunks.inc"
#undef TVM_AIR_CHUNK

// d_main: [>=379][r*n] coset-major LDE of the main table; d_aux: [>=270][r*n] (X-field columns as
// 3 planar B-field columns); d_challenges: 63 X-field (Montgomery); d_weights: 604 X-field;
// d_out: 3 planes of r*n words (memory order = coset-major), overwritten.
void air_quotient_run(Ctx &c, const u64 *d_main, size_t main_stride, const u64 *d_aux, size_t aux_stride,
                      const u64 *d_challenges, const u64 *d_weights, unsigned log_n, unsigned log_r,
                      u64 offset_mont, u64 *d_out, size_t out_stride) {
  if ((1u << log_r) > (unsigned)AIR_MAX_COSETS) throw ApiError{TVM_ERR_UNSUPPORTED, "too many cosets"};
  AirArgs a{};
  a.main = d_main; a.main_stride = main_stride; a.aux = d_aux; a.aux_stride = aux_stride;
  a.out = d_out; a.out_stride = out_stride;
  a.nrows = (size_t)1 << (log_n + log_r);
  a.log_n = (int)log_n;
  u64 wn = root_of_unity_mont(log_n);
  a.trace_gen = c.get_pow_tab(wn, (int)log_n);
  a.trace_gen_inv = finv(wn);
  u64 wrn = root_of_unity_mont(log_n + log_r);
  u64 x = offset_mont;
  for (unsigned cs = 0; cs < (1u << log_r); cs++) {
    a.coset_x[cs] = x;
    a.cons_zerofier_inv[cs] = finv(fsub(fpow(x, (u64)1 << log_n), MONT_ONE));
    x = fmul(x, wrn);
  }
  TVM_CUDA(cudaMemsetAsync(d_out, 0, sizeof(u64) * a.nrows, c.stream));
  TVM_CUDA(cudaMemsetAsync(d_out + out_stride, 0, sizeof(u64) * a.nrows, c.stream));
  TVM_CUDA(cudaMemsetAsync(d_out + 2 * out_stride, 0, sizeof(u64) * a.nrows, c.stream));
#define TVM_AIR_CHUNK(name) name##_launch(a, d_weights, d_challenges, c.stream); c.launches++;
#include "air_gen/air_chunks.inc"
#undef TVM_AIR_CHUNK
  TVM_CUDA(cudaGetLastError());
}

}  // namespace tvm

// Shared declarations for the generated AIR quotient kernels (csrc/air_gen/*.cu).
// Semantics: triton-vm/src/table/master_table.rs:1194-1363 (zerofier inverses +
// all_quotients_combined).  Row layout: tables are column-major and coset-major, memory index
// m = coset * n + k  <->  quotient-domain row i = coset + r*k, x_i = offset * w_{rn}^i; the
// "next row" i + r (unit distance = quotient_len / trace_len, master_table.rs:1305-1306) is
// (coset, k+1 mod n).
#pragma once
#include "ctx.h"

namespace tvm {

#ifndef TVM_AIR_THREADS
#define TVM_AIR_THREADS 256
#endif
static constexpr int AIR_THREADS = TVM_AIR_THREADS;
static constexpr int AIR_MAX_COSETS = 64;

struct AirArgs {
  const u64 *main;      // main column q at main + q*main_stride
  size_t main_stride;
  const u64 *aux;       // aux X-field column q, coordinate d at aux + (3q+d)*aux_stride
  size_t aux_stride;
  u64 *out;             // quotient codeword, planar: coordinate d at out + d*out_stride, memory order
  size_t out_stride;
  size_t nrows;         // r * n
  int log_n;            // trace length n = 2^log_n
  PowTab trace_gen;     // w_n^k
  u64 trace_gen_inv;    // w_n^-1
  u64 coset_x[AIR_MAX_COSETS];            // offset * w_{rn}^coset
  u64 cons_zerofier_inv[AIR_MAX_COSETS];  // 1 / (x^n - 1), constant on a coset
};

__device__ __forceinline__ u64 air_domain_value(const AirArgs &a, size_t coset, size_t k) {
  u64 lo = __ldg(a.trace_gen.lo + (k & ((1ULL << a.trace_gen.shift) - 1)));
  u64 hi = __ldg(a.trace_gen.hi + (k >> a.trace_gen.shift));
  return fmul(a.coset_x[coset], fmul(lo, hi));
}

__device__ __forceinline__ void air_add_out(const AirArgs &a, size_t m, xfe v) {
  u64 *o = a.out + m;
  o[0] = fadd(o[0], v.c0);
  o[a.out_stride] = fadd(o[a.out_stride], v.c1);
  o[2 * a.out_stride] = fadd(o[2 * a.out_stride], v.c2);
}

// initial: zerofier x - 1 (master_table.rs:1194-1202)
__device__ __forceinline__ void air_accumulate_init(const AirArgs &a, size_t m, size_t coset, size_t k, xfe acc) {
  u64 x = air_domain_value(a, coset, k);
  air_add_out(a, m, xmulb(acc, finv(fsub(x, MONT_ONE))));
}
// consistency: zerofier x^n - 1 (1204-1214)
__device__ __forceinline__ void air_accumulate_cons(const AirArgs &a, size_t m, size_t coset, size_t, xfe acc) {
  air_add_out(a, m, xmulb(acc, a.cons_zerofier_inv[coset]));
}
// transition: zerofier (x^n - 1) / (x - w_n^-1) (1216-1237)
__device__ __forceinline__ void air_accumulate_tran(const AirArgs &a, size_t m, size_t coset, size_t k, xfe acc) {
  u64 x = air_domain_value(a, coset, k);
  air_add_out(a, m, xmulb(acc, fmul(fsub(x, a.trace_gen_inv), a.cons_zerofier_inv[coset])));
}
// terminal: zerofier x - w_n^-1 (1239-1252)
__device__ __forceinline__ void air_accumulate_term(const AirArgs &a, size_t m, size_t coset, size_t k, xfe acc) {
  u64 x = air_domain_value(a, coset, k);
  air_add_out(a, m, xmulb(acc, finv(fsub(x, a.trace_gen_inv))));
}

}  // namespace tvm

// MasterMainTable::new + pad from the AlgebraicExecutionTrace: the nine tables' `fill` and `pad`
// (reference triton-vm/src/table/master_table.rs:881-974 and the per-table impls cited at each stage), as data-parallel
// stages over an executor `Ex`.
//
// The stages are written once, as element bodies plus the orchestration around them, against a small executor interface:
//     u64 *alloc(size_t words)                          device words, released with the executor
//     void launch(size_t count, F body)                 body(i) for every i < count
//     void sort_perm(const u64 *keys, u64 *perm, n)     the STABLE sorting permutation of the keys (perm[k] = source index)
//     void exclusive_sum(const u64 *in, u64 *out, n)
//     void ntt(const u64 *in, u64 *out, log_n, count, inverse)   `count` contiguous natural-order transforms (Montgomery; the
//                                                       inverse includes 1/n), in != out
//     u64 read_word(const u64 *p)                       one word back to the host (synchronises)
// The library instantiates them with the CUDA executor of main_fill.cu (kernels, cub::DeviceRadixSort / DeviceScan, ntt_run).
// tests/host/main_fill_host.cu instantiates the same text with a sequential host executor so that the host-only test suite
// can check the logic against the oracle's table fill without a GPU; that executor is test infrastructure and is not
// part of the library.
//
// Representation: the AET arrays and the table are CANONICAL words (the tables are sorted and compared by `.value()`);
// field arithmetic (inverses, the Bezout polynomials) converts to Montgomery locally.
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <vector>
#include "../field.cuh"

namespace tvm {
namespace fill {

#define TVM_FILL_BODY [=] __host__ __device__

// ---- main-table layout (triton-air/src/table_column.rs; airgen/columns.py prints the same offsets) ----
enum : unsigned {
  COL_PROGRAM = 0, COL_PROCESSOR = 7, COL_OP_STACK = 46, COL_RAM = 50, COL_JUMP_STACK = 57, COL_HASH = 62, COL_CASCADE = 129,
  COL_LOOKUP = 135, COL_U32 = 139, NUM_TABLE_COLUMNS = 149,
  W_PROCESSOR = 39, W_OP_STACK = 4, W_RAM = 7, W_HASH = 67,
  // processor-trace columns used by other tables
  PROC_CLK = 0, PROC_IS_PADDING = 1, PROC_CI = 3, PROC_JSP = 12, PROC_JSO = 13, PROC_JSD = 14, PROC_CJD_MULT = 38,
  // opcodes (triton-isa/src/instruction.rs:315-364)
  OP_SPLIT = 4, OP_LT = 6, OP_LOG2FLOOR = 12, OP_AND = 14, OP_HASH = 18, OP_POP_COUNT = 28, OP_POW = 30,
  // constant block (FillConsts, uploaded once per fill)
  C_INV9 = 0,          // inverse_or_zero(9 - k), k < 10            (program.rs:33-113: MaxMinusIndexInChunkInv)
  C_BITS33 = 10,       // (k - 33)^-1, k < 34                       (u32.rs:111, 236-237)
  C_RC0 = 44,          // Tip5 round constants of round 0 (canonical) (hash.rs:283-291)
  C_HASH_INV0 = 60,    // inverse_or_zero(2^32 - 1 - 0)             (hash.rs:270: inverse_or_zero_of_highest_2_limbs(0))
  C_LUT = 61,          // Tip5 lookup table, 256 entries            (cascade.rs:29-35, lookup.rs:84-116)
  C_WORDS = 61 + 256
};

struct AetView {   // aet.rs:41-91; every pointer addresses canonical words in the executor's memory space
  const u64 *program; size_t program_len;              // Program::to_bwords()
  const u64 *instruction_multiplicities;               // [program_len]
  const u64 *processor_trace; size_t processor_rows;   // [rows][39]
  const u64 *op_stack_trace; size_t op_stack_rows;     // [rows][4]
  const u64 *ram_trace; size_t ram_rows;               // [rows][7]
  const u64 *program_hash_trace; size_t program_hash_rows;   // [rows][67]
  const u64 *sponge_trace; size_t sponge_rows;         // [rows][67]
  const u64 *hash_trace; size_t hash_rows;             // [rows][67]
  const u64 *u32_entries; size_t u32_count;            // [count][4]: opcode, lhs, rhs, multiplicity (IndexMap order)
  const u64 *cascade_multiplicities; size_t cascade_count;   // [count][2]: 16-bit limb, multiplicity (IndexMap order)
  const u64 *lookup_multiplicities;                    // [256]
};

// ---- small helpers usable on both sides --------------------------------------------------------------
TVM_HD u64 inv_or_zero_canon(u64 x) { return x ? from_mont(finv(to_mont(x))) : 0; }
TVM_HD u64 mul_canon(u64 a, u64 b) { return fmul(to_mont(a), b); }          // (aR)(b)/R = ab
TVM_HD void count_add(u64 *p, u64 v) {
#ifdef __CUDA_ARCH__
  atomicAdd((unsigned long long *)p, (unsigned long long)v);
#else
  *p += v;   // the host executor runs bodies sequentially
#endif
}
inline unsigned ceil_log2(size_t x) { unsigned l = 0; while (((size_t)1 << l) < x) l++; return l; }

// Host-side constants of the fill (canonical); `rc0_canon` = Tip5 round constants 0..15, `lut` = Tip5's 8-bit lookup table.
inline std::vector<u64> fill_constants(const u64 *rc0_canon, const unsigned char *lut) {
  std::vector<u64> c(C_WORDS, 0);
  for (unsigned k = 0; k < 10; k++) c[C_INV9 + k] = inv_or_zero_canon(9 - k);
  for (unsigned k = 0; k < 34; k++) c[C_BITS33 + k] = inv_or_zero_canon(fsub((u64)k, 33));
  for (unsigned k = 0; k < 16; k++) c[C_RC0 + k] = rc0_canon[k];
  c[C_HASH_INV0] = inv_or_zero_canon(0xFFFFFFFFULL);
  for (unsigned k = 0; k < 256; k++) c[C_LUT + k] = lut[k];
  return c;
}

// =====================================================================================================
// Bezout coefficient polynomials of rp = prod (x - r_i) and rp'  (ram.rs:162-214).
//   b interpolates 1/rp'(r_i) in the roots, a = (1 - rp' b)/rp; the reference does this with twenty-first's fast
//   zerofier / batch evaluation / interpolation.  Here: one subproduct tree, three passes over it, all batched
//   transforms:
//     up    P_v = P_left P_right                                       (monic, stored without the leading coefficient)
//     down  rp'(r_i) through scaled remainders: the first c_v coefficients of (rp' mod P_v)/P_v as a series in 1/x go to a
//           child by a middle product with the sibling's polynomial; at the root they are the power sums of the roots,
//           rev(rp')/rev(rp) mod y^m (Newton iteration for the reciprocal); at a leaf the first coefficient is rp'(r_i)
//     up    b = sum_i w_i rp/(x - r_i), w_i = rp'(r_i)^-2:  B_v = B_left P_right + B_right P_left
//     a     by evaluation on a coset of >= m+1 points where rp has no root
//   Nodes below `DIRECT` coefficients use schoolbook products inside one body; larger nodes use transforms of twice their
//   size, and the transforms of the node polynomials are kept from the first pass for the other two.
//   All values Montgomery.  roots: m >= 1 distinct elements.  a_out, b_out: m words each.
// =====================================================================================================
static constexpr unsigned BEZOUT_DIRECT_LOG = 5;   // nodes of up to 2^5 coefficients: schoolbook

struct TreeShape {
  size_t m; unsigned L; size_t M;
  TVM_HD size_t count(unsigned level, size_t j) const {   // real roots under node j of `level`
    const size_t d = (size_t)1 << level, lo = j << level;
    return lo >= m ? 0 : (m - lo < d ? m - lo : d);
  }
};

// out = a * b, full product of lengths la, lb (out has la + lb - 1 words); sizes decide between one body per output
// coefficient and transforms
template <class Ex>
void poly_mul(Ex &ex, const u64 *a, size_t la, const u64 *b, size_t lb, u64 *out) {
  const size_t lo = la + lb - 1;
  if ((la < lb ? la : lb) <= 32) {
    ex.launch(lo, TVM_FILL_BODY(size_t k) {
      u64 acc = 0;
      const size_t i0 = k + 1 > lb ? k + 1 - lb : 0, i1 = k < la - 1 ? k : la - 1;
      for (size_t i = i0; i <= i1; i++) acc = fadd(acc, fmul(a[i], b[k - i]));
      out[k] = acc;
    });
    return;
  }
  const unsigned lg = ceil_log2(lo);
  const size_t N = (size_t)1 << lg;
  u64 *buf = ex.alloc(4 * N);
  u64 *pa = buf, *pb = buf + N, *fa = buf + 2 * N, *fb = buf + 3 * N;
  ex.launch(N, TVM_FILL_BODY(size_t k) { pa[k] = k < la ? a[k] : 0; pb[k] = k < lb ? b[k] : 0; });
  ex.ntt(pa, fa, lg, 2, false);          // pa, pb are contiguous: one batched call
  ex.launch(N, TVM_FILL_BODY(size_t k) { fa[k] = fmul(fa[k], fb[k]); });
  ex.ntt(fa, pa, lg, 1, true);
  ex.launch(lo, TVM_FILL_BODY(size_t k) { out[k] = pa[k]; });
}

// transforms (size 2^(l+1)) of the M >> l node polynomials of level l, each zero-padded to twice its size; pad, out: 2M words
template <class Ex>
void node_transforms(Ex &ex, const u64 *src, unsigned l, size_t M, u64 *pad, u64 *out) {
  const size_t d = (size_t)1 << l;
  ex.launch(2 * M, TVM_FILL_BODY(size_t i) {
    const size_t node = i >> (l + 1), k = i & (2 * d - 1);
    pad[i] = k < d ? src[node * d + k] : 0;
  });
  ex.ntt(pad, out, l + 1, M >> l, false);
}

template <class Ex>
void bezout_coefficients(Ex &ex, const u64 *roots, size_t m, u64 *a_out, u64 *b_out, unsigned direct_log = BEZOUT_DIRECT_LOG) {
  if (m == 0) return;
  TreeShape sh{m, ceil_log2(m), (size_t)1 << ceil_log2(m)};
  const unsigned L = sh.L;
  const size_t M = sh.M;
  const unsigned D = direct_log;      // (tests lower it to reach the transform paths with few roots)

  // ---- pass 1: subproduct tree.  low[l]: node j at [j 2^l, (j+1) 2^l): P_v - x^{c_v};  hat[l] (l >= D): its transform of
  //      size 2^(l+1) at [j 2^(l+1), ...)
  std::vector<u64 *> low(L + 1), hat(L + 1, nullptr);
  for (unsigned l = 0; l <= L; l++) low[l] = ex.alloc(M);
  {
    u64 *l0 = low[0];
    ex.launch(M, TVM_FILL_BODY(size_t j) { l0[j] = j < m ? fneg(roots[j]) : 0; });
  }
  u64 *pad = ex.alloc(2 * M), *prod = ex.alloc(2 * M), *prod_t = ex.alloc(2 * M);
  auto make_hat = [&](unsigned l) {       // transforms of all nodes of level l, zero-padded to twice their size
    if (hat[l]) return;
    hat[l] = ex.alloc(2 * M);
    node_transforms(ex, low[l], l, M, pad, hat[l]);
  };
  for (unsigned l = 0; l < L; l++) {      // level l -> l + 1
    const size_t d = (size_t)1 << l;
    const u64 *src = low[l];
    u64 *dst = low[l + 1];
    const u64 *cyc = nullptr;             // cyclic products A B of sibling pairs, size 2d each
    if (l >= D) {
      make_hat(l);
      const u64 *h = hat[l];
      u64 *pt = prod_t;
      ex.launch(M, TVM_FILL_BODY(size_t i) {      // pair q = i / 2d: nodes 2q, 2q+1
        const size_t q = i >> (l + 1), k = i & (2 * d - 1);
        pt[i] = fmul(h[(2 * q) * 2 * d + k], h[(2 * q + 1) * 2 * d + k]);
      });
      ex.ntt(prod_t, prod, l + 1, M >> (l + 1), true);
      cyc = prod;
    }
    ex.launch(M, TVM_FILL_BODY(size_t i) {
      const size_t q = i >> (l + 1), k = i & (2 * d - 1);
      const u64 *A = src + (2 * q) * d, *B = src + (2 * q + 1) * d;
      const size_t cA = sh.count(l, 2 * q), cB = sh.count(l, 2 * q + 1);
      u64 acc;
      if (cyc) acc = cyc[i];
      else {
        acc = 0;
        const size_t i0 = k + 1 > d ? k + 1 - d : 0, i1 = k < d - 1 ? k : d - 1;
        for (size_t t = i0; t <= i1; t++) acc = fadd(acc, fmul(A[t], B[k - t]));
      }
      if (k >= cB && k - cB < d) acc = fadd(acc, A[k - cB]);     // A x^cB
      if (k >= cA && k - cA < d) acc = fadd(acc, B[k - cA]);     // B x^cA
      dst[i] = acc;                                              // (the leading x^(cA+cB) is implicit)
    });
  }
  const u64 *rp_low = low[L];             // rp = x^m + rp_low[0..m)

  // ---- root series: rho_k = coefficient of y^k in rev(rp')/rev(rp), k < m  (the power sums of the roots)
  u64 *rho = ex.alloc(M > 1 ? M : 2);
  {
    size_t P2 = 1;
    while (P2 < m) P2 <<= 1;
    u64 *h = ex.alloc(2 * P2 + 2), *g = ex.alloc(2 * P2 + 2), *e = ex.alloc(4 * P2 + 4), *g2 = ex.alloc(4 * P2 + 4);
    ex.launch(2 * P2, TVM_FILL_BODY(size_t k) {      // h = rev(rp) = 1 + sum_{k=1..m} rp_{m-k} y^k
      h[k] = k == 0 ? MONT_ONE : (k <= m ? rp_low[m - k] : 0);
      g[k] = k == 0 ? MONT_ONE : 0;
    });
    for (size_t k = 1; k < P2; k <<= 1) {            // g <- g (2 - h g) mod y^2k
      poly_mul(ex, h, 2 * k, g, k, e);               // 3k - 1 coefficients
      ex.launch(2 * k, TVM_FILL_BODY(size_t i) { e[i] = i == 0 ? fsub(fadd(MONT_ONE, MONT_ONE), e[0]) : fneg(e[i]); });
      poly_mul(ex, e, 2 * k, g, k, g2);
      ex.launch(2 * k, TVM_FILL_BODY(size_t i) { g[i] = g2[i]; });
    }
    // rev(rp')_k = (m - k) rp_{m-k}, k < m  (rp_m = 1)
    u64 *rf = e;
    ex.launch(m, TVM_FILL_BODY(size_t k) {
      const u64 c = k == 0 ? MONT_ONE : rp_low[m - k];
      rf[k] = fmul(to_mont((u64)(m - k)), c);
    });
    poly_mul(ex, rf, m, g, m, g2);
    ex.launch(M, TVM_FILL_BODY(size_t k) { rho[k] = k < m ? g2[k] : 0; });
  }

  // ---- pass 2: down.  sig: node j of level l at [j 2^l, ...) holds the first c_v series coefficients
  u64 *sig = rho, *sig_next = ex.alloc(M > 1 ? M : 2);
  for (unsigned l = L; l > 0; l--) {      // level l -> l - 1; d = child size
    const size_t d = (size_t)1 << (l - 1);
    const u64 *child_low = low[l - 1];
    const u64 *s = sig;
    u64 *out = sig_next;
    const u64 *corr = nullptr;            // per parent: [corr(B_right, s) | corr(A_left, s)], 2d words each
    if (l - 1 >= D) {
      make_hat(l - 1);
      const u64 *h = hat[l - 1];
      ex.ntt(sig, prod_t, l, M >> l, false);          // transforms of the parents' series, size 2d
      u64 *tmp = pad;                     // 2M words: two products per parent
      const u64 *st = prod_t;
      ex.launch(2 * M, TVM_FILL_BODY(size_t i) {
        const size_t q = i >> (l + 1), which = (i >> l) & 1, k = i & (2 * d - 1);
        // child `which` (0 = left) is fed by its sibling's polynomial, transform index -k
        const size_t sib = 2 * q + (1 - which);
        const size_t kk = (2 * d - k) & (2 * d - 1);
        tmp[i] = fmul(st[q * 2 * d + k], h[sib * 2 * d + kk]);
      });
      ex.ntt(pad, prod, l, M >> (l - 1), true);       // 2 transforms of size 2d per parent
      corr = prod;
    }
    ex.launch(M, TVM_FILL_BODY(size_t i) {            // i = child node * d + t
      const size_t node = i >> (l - 1), t = i & (d - 1), q = node >> 1, which = node & 1;
      const size_t sib = node ^ 1;
      const size_t c_self = sh.count(l - 1, node), c_sib = sh.count(l - 1, sib);
      u64 acc = 0;
      if (t < c_self) {
        const u64 *sp = s + q * 2 * d;
        if (corr) acc = corr[(2 * q + which) * 2 * d + t];
        else {
          const u64 *S = child_low + sib * d;
          for (size_t j = 0; j < c_sib; j++) acc = fadd(acc, fmul(S[j], sp[t + j]));
        }
        acc = fadd(acc, sp[t + c_sib]);               // the sibling's leading coefficient
      }
      out[i] = acc;
    });
    u64 *sw = sig; sig = sig_next; sig_next = sw;
  }
  // sig[i] = rp'(r_i).  w_i = rp'(r_i)^-2
  u64 *bcur = ex.alloc(M > 1 ? M : 2), *bnext = sig_next;
  {
    const u64 *s = sig;
    u64 *w = bcur;
    ex.launch(M, TVM_FILL_BODY(size_t i) {
      u64 v = 0;
      if (i < m) { const u64 f = finv(s[i]); v = fmul(f, f); }
      w[i] = v;
    });
  }
  // ---- pass 3: up.  B_v = B_left P_right + B_right P_left
  for (unsigned l = 0; l < L; l++) {
    const size_t d = (size_t)1 << l;
    const u64 *src = bcur, *plow = low[l];
    u64 *dst = bnext;
    const u64 *cyc = nullptr;
    if (l >= D) {
      const u64 *h = hat[l];
      u64 *p = pad;
      ex.launch(2 * M, TVM_FILL_BODY(size_t i) {
        const size_t node = i >> (l + 1), k = i & (2 * d - 1);
        p[i] = k < d ? src[node * d + k] : 0;
      });
      ex.ntt(pad, prod, l + 1, M >> l, false);
      const u64 *bh = prod;
      u64 *pt = prod_t;
      ex.launch(M, TVM_FILL_BODY(size_t i) {
        const size_t q = i >> (l + 1), k = i & (2 * d - 1);
        const size_t a = (2 * q) * 2 * d + k, b = (2 * q + 1) * 2 * d + k;
        pt[i] = fadd(fmul(bh[a], h[b]), fmul(bh[b], h[a]));
      });
      ex.ntt(prod_t, pad, l + 1, M >> (l + 1), true);
      cyc = pad;
    }
    ex.launch(M, TVM_FILL_BODY(size_t i) {
      const size_t q = i >> (l + 1), k = i & (2 * d - 1);
      const u64 *BL = src + (2 * q) * d, *BR = src + (2 * q + 1) * d;
      const u64 *PL = plow + (2 * q) * d, *PR = plow + (2 * q + 1) * d;
      const size_t cL = sh.count(l, 2 * q), cR = sh.count(l, 2 * q + 1);
      u64 acc;
      if (cyc) acc = cyc[i];
      else {
        acc = 0;
        const size_t i0 = k + 1 > d ? k + 1 - d : 0, i1 = k < d - 1 ? k : d - 1;
        for (size_t t = i0; t <= i1; t++) acc = fadd(acc, fadd(fmul(BL[t], PR[k - t]), fmul(BR[t], PL[k - t])));
      }
      if (k >= cR && k - cR < d) acc = fadd(acc, BL[k - cR]);
      if (k >= cL && k - cL < d) acc = fadd(acc, BR[k - cL]);
      dst[i] = acc;
    });
    u64 *sw = bcur; bcur = bnext; bnext = sw;
  }
  {
    const u64 *b = bcur;
    ex.launch(m, TVM_FILL_BODY(size_t k) { b_out[k] = b[k]; });
  }

  // ---- a = (1 - rp' b) / rp on a coset g <w_N>, N >= m + 1 points, g = 7^j for the first j with no root of rp on it
  {
    const unsigned lg = ceil_log2(m + 1);
    const size_t N = (size_t)1 << lg;
    u64 *in3 = ex.alloc(3 * N), *ev3 = ex.alloc(3 * N), *flag = ex.alloc(1);
    const u64 *b = bcur;
    bool done = false;
    u64 g = to_mont(7);
    for (int attempt = 0; attempt < 16 && !done; attempt++, g = fmul(g, to_mont(7))) {
      ex.launch(1, TVM_FILL_BODY(size_t) { flag[0] = 0; });
      ex.launch(N, TVM_FILL_BODY(size_t k) {
        const u64 gk = fpow(g, (u64)k);
        const u64 rp_k = k < m ? rp_low[k] : (k == m ? MONT_ONE : 0);
        const u64 fd_k = k < m ? fmul(to_mont((u64)(k + 1)), k + 1 < m ? rp_low[k + 1] : MONT_ONE) : 0;
        in3[k] = fmul(rp_k, gk);
        in3[N + k] = fmul(fd_k, gk);
        in3[2 * N + k] = k < m ? fmul(b[k], gk) : 0;
      });
      ex.ntt(in3, ev3, lg, 3, false);
      ex.launch(N, TVM_FILL_BODY(size_t k) {
        const u64 den = ev3[k];
        if (den == 0) { count_add(flag, 1); return; }
        ev3[k] = fmul(fsub(MONT_ONE, fmul(ev3[N + k], ev3[2 * N + k])), finv(den));
      });
      if (ex.read_word(flag) != 0) continue;
      ex.ntt(ev3, in3, lg, 1, true);
      const u64 ginv = finv(g);
      ex.launch(m, TVM_FILL_BODY(size_t k) { a_out[k] = fmul(in3[k], fpow(ginv, (u64)k)); });
      done = true;
    }
    if (!done) throw std::runtime_error("bezout_coefficients: no coset without a root of the RAM-pointer polynomial");
  }
}

// =====================================================================================================
// The nine tables.  T: [149][n] canonical, column c at T + c*n.  n = number of rows of the (padded) trace table.
// =====================================================================================================
struct FillInfo {       // what MasterMainTable::new records about the AET (master_table.rs:890-897)
  size_t program_len_padded, processor_len, op_stack_len, ram_len, hash_len, cascade_len, u32_len, unique_ram_pointers;
};

// u32.rs:193-290 `u32_section_next_row`: number of rows of the section of one entry
TVM_HD unsigned u32_section_len(u64 op, u64 lhs, u64 rhs) {
  unsigned k = 0;
  while (k < 40) {
    if ((lhs == 0 || op == OP_POW) && rhs == 0) break;
    if (op != OP_POW) lhs >>= 1;
    rhs >>= 1;
    k++;
  }
  return k + 1;
}

template <class Ex>
FillInfo main_table_from_aet(Ex &ex, const AetView &aet, const u64 *consts /* fill_constants(), executor memory */, size_t n, u64 *T,
                             unsigned bezout_direct_log = BEZOUT_DIRECT_LOG) {
  const size_t plen = aet.processor_rows;
  if (plen == 0) throw std::invalid_argument("the processor trace must have at least one row");
  const size_t prog_len = aet.program_len, prog_padded = (prog_len + 1 + 9) / 10 * 10;
  const size_t hash_len = aet.program_hash_rows + aet.sponge_rows + aet.hash_rows;
  FillInfo info{prog_padded, plen, aet.op_stack_rows, aet.ram_rows, hash_len, aet.cascade_count, 0, 0};
  if (n < 2 || (n & (n - 1)) || n < plen || n < prog_padded || n < aet.op_stack_rows || n < aet.ram_rows || n < hash_len ||
      n < aet.cascade_count || n < 256)
    throw std::invalid_argument("the table height is below a table's length (or not a power of two)");

  // ---- program table (program.rs:33-113)
  {
    u64 *t = T + (size_t)COL_PROGRAM * n;
    const u64 *program = aet.program, *mult = aet.instruction_multiplicities;
    ex.launch(n, TVM_FILL_BODY(size_t i) {
      t[0 * n + i] = i;                                                   // Address
      t[1 * n + i] = i < prog_len ? program[i] : (i == prog_len ? 1 : 0); // Instruction: program, then 1 0 0 .. (aet.rs:214-226)
      t[2 * n + i] = i < prog_len ? mult[i] : 0;                          // LookupMultiplicity
      t[3 * n + i] = i % 10;                                              // IndexInChunk
      t[4 * n + i] = consts[C_INV9 + i % 10];                             // MaxMinusIndexInChunkInv
      t[5 * n + i] = i >= prog_len;                                       // IsHashInputPadding
      t[6 * n + i] = i >= prog_padded;                                    // IsTablePadding
    });
  }

  // clock-jump-difference multiplicities (processor.rs:50-61): hist[clk difference], filled by the three memory-like tables
  u64 *hist = ex.alloc(plen + 1);
  ex.launch(plen + 1, TVM_FILL_BODY(size_t i) { hist[i] = 0; });

  // ---- op-stack table (op_stack.rs:179-211): sorted by (stack pointer, clk); the trace is in clk order, the sort stable
  {
    const size_t len = aet.op_stack_rows;
    const u64 *tr = aet.op_stack_trace;
    u64 *t = T + (size_t)COL_OP_STACK * n;
    u64 *perm = ex.alloc(len + 1);
    if (len) {
      u64 *keys = ex.alloc(len);
      ex.launch(len, TVM_FILL_BODY(size_t i) { keys[i] = tr[i * W_OP_STACK + 2]; });
      ex.sort_perm(keys, perm, len);
      ex.launch(len - 1, TVM_FILL_BODY(size_t i) {
        const u64 *a = tr + perm[i] * W_OP_STACK, *b = tr + perm[i + 1] * W_OP_STACK;
        if (a[2] == b[2]) { const u64 dclk = b[0] - a[0]; if (dclk <= plen) count_add(hist + dclk, 1); }
      });
    }
    ex.launch(n, TVM_FILL_BODY(size_t i) {
      if (len == 0) {                                                     // pad of an empty table (op_stack.rs:197-211)
        t[0 * n + i] = 0; t[1 * n + i] = 2; t[2 * n + i] = 16; t[3 * n + i] = 0;
        return;
      }
      const u64 *r = tr + perm[i < len ? i : len - 1] * W_OP_STACK;
      t[0 * n + i] = r[0];
      t[1 * n + i] = i < len ? r[1] : 2;                                  // IB1ShrinkStack = PADDING_VALUE
      t[2 * n + i] = r[2];
      t[3 * n + i] = r[3];
    });
  }

  // ---- RAM table (ram.rs:64-103, 216-262): sorted by (pointer, clk); inverse of the pointer difference; the Bezout
  //      coefficients are handed out from the highest one downwards, a new one whenever the pointer changes
  {
    const size_t len = aet.ram_rows;
    const u64 *tr = aet.ram_trace;
    u64 *t = T + (size_t)COL_RAM * n;
    u64 *perm = ex.alloc(len + 1), *rank = ex.alloc(len + 1), *bez_a = nullptr, *bez_b = nullptr;
    size_t m = 0;
    if (len) {
      u64 *keys = ex.alloc(len), *flags = ex.alloc(len);
      ex.launch(len, TVM_FILL_BODY(size_t i) { keys[i] = tr[i * W_RAM + 2]; });
      ex.sort_perm(keys, perm, len);
      ex.launch(len, TVM_FILL_BODY(size_t i) { flags[i] = i > 0 && tr[perm[i] * W_RAM + 2] != tr[perm[i - 1] * W_RAM + 2]; });
      ex.exclusive_sum(flags, rank, len);
      ex.launch(len, TVM_FILL_BODY(size_t i) { rank[i] += flags[i]; });   // index of the row's pointer among the unique ones
      m = (size_t)ex.read_word(rank + (len - 1)) + 1;
      u64 *roots = ex.alloc(m);
      bez_a = ex.alloc(m); bez_b = ex.alloc(m);
      ex.launch(len, TVM_FILL_BODY(size_t i) { if (i == 0 || flags[i]) roots[rank[i]] = to_mont(tr[perm[i] * W_RAM + 2]); });
      bezout_coefficients(ex, roots, m, bez_a, bez_b, bezout_direct_log);
      ex.launch(len - 1, TVM_FILL_BODY(size_t i) {
        if (flags[i + 1]) return;
        const u64 dclk = tr[perm[i + 1] * W_RAM] - tr[perm[i] * W_RAM];
        if (dclk <= plen) count_add(hist + dclk, 1);
      });
    }
    info.unique_ram_pointers = m;
    ex.launch(n, TVM_FILL_BODY(size_t i) {
      if (len == 0) {                                                     // pad of an empty table (ram.rs:88-103)
        for (unsigned c = 0; c < W_RAM; c++) t[c * n + i] = 0;
        t[1 * n + i] = 2; t[6 * n + i] = 1;
        return;
      }
      const size_t k = i < len ? i : len - 1;
      const u64 *r = tr + perm[k] * W_RAM;
      t[0 * n + i] = r[0];
      t[1 * n + i] = i < len ? r[1] : 2;                                  // InstructionType = PADDING_INDICATOR
      t[2 * n + i] = r[2];
      t[3 * n + i] = r[3];
      t[4 * n + i] = k + 1 < len ? inv_or_zero_canon(fsub(tr[perm[k + 1] * W_RAM + 2], r[2])) : 0;
      t[5 * n + i] = from_mont(bez_a[m - 1 - rank[k]]);
      t[6 * n + i] = from_mont(bez_b[m - 1 - rank[k]]);
    });
  }

  // ---- jump-stack table (jump_stack.rs:90-205): the processor's rows grouped by JSP (stable); the padding rows follow the
  //      row of the last clock cycle, the rows after it move to the end
  {
    const u64 *tr = aet.processor_trace;
    u64 *t = T + (size_t)COL_JUMP_STACK * n;
    u64 *perm = ex.alloc(plen), *keys = ex.alloc(plen), *kmax = ex.alloc(1);
    ex.launch(plen, TVM_FILL_BODY(size_t i) { keys[i] = tr[i * W_PROCESSOR + PROC_JSP]; });
    ex.sort_perm(keys, perm, plen);
    ex.launch(1, TVM_FILL_BODY(size_t) { kmax[0] = plen - 1; });      // (a trace whose CLK column is not the row index)
    ex.launch(plen, TVM_FILL_BODY(size_t i) {
      if (tr[perm[i] * W_PROCESSOR + PROC_CLK] == plen - 1) kmax[0] = i;
      if (i + 1 < plen) {
        const u64 *a = tr + perm[i] * W_PROCESSOR, *b = tr + perm[i + 1] * W_PROCESSOR;
        if (a[PROC_JSP] == b[PROC_JSP]) { const u64 dclk = b[PROC_CLK] - a[PROC_CLK]; if (dclk <= plen) count_add(hist + dclk, 1); }
      }
    });
    const size_t npad = n - plen;
    ex.launch(n, TVM_FILL_BODY(size_t i) {
      const size_t km = (size_t)kmax[0];
      size_t src;
      u64 clk;
      if (i <= km) { src = perm[i]; clk = tr[src * W_PROCESSOR + PROC_CLK]; }
      else if (i < km + 1 + npad) { src = perm[km]; clk = plen + (i - km - 1); }
      else { src = perm[i - npad]; clk = tr[src * W_PROCESSOR + PROC_CLK]; }
      const u64 *r = tr + src * W_PROCESSOR;
      t[0 * n + i] = clk;
      t[1 * n + i] = r[PROC_CI];
      t[2 * n + i] = r[PROC_JSP];
      t[3 * n + i] = r[PROC_JSO];
      t[4 * n + i] = r[PROC_JSD];
    });
  }

  // ---- processor table (processor.rs:45-95)
  {
    const u64 *tr = aet.processor_trace;
    u64 *t = T + (size_t)COL_PROCESSOR * n;
    ex.launch(n, TVM_FILL_BODY(size_t i) {
      const u64 *r = tr + (i < plen ? i : plen - 1) * W_PROCESSOR;
      for (unsigned c = 0; c < W_PROCESSOR; c++) t[c * n + i] = r[c];
      if (i >= plen) { t[PROC_CLK * n + i] = i; t[PROC_IS_PADDING * n + i] = 1; }
      // one lookup of clock jump difference 1 per padding row of the jump-stack table (processor.rs:85-94)
      t[PROC_CJD_MULT * n + i] = (i < plen ? hist[i] : 0) + (i == 1 ? n - plen : 0);
    });
  }

  // ---- hash table (hash.rs:241-302): program hashing, sponge, hash sections, then padding
  {
    const u64 *ph = aet.program_hash_trace, *sp = aet.sponge_trace, *hs = aet.hash_trace;
    const size_t n_ph = aet.program_hash_rows, n_sp = aet.sponge_rows;
    u64 *t = T + (size_t)COL_HASH * n;
    ex.launch(n, TVM_FILL_BODY(size_t i) {
      if (i < hash_len) {
        const u64 *r;
        u64 mode;
        if (i < n_ph) { r = ph + i * W_HASH; mode = 1; }
        else if (i < n_ph + n_sp) { r = sp + (i - n_ph) * W_HASH; mode = 2; }
        else { r = hs + (i - n_ph - n_sp) * W_HASH; mode = 3; }
        for (unsigned c = 1; c < W_HASH; c++) t[c * n + i] = r[c];
        t[i] = mode;
      } else {
        for (unsigned c = 0; c < W_HASH; c++) t[c * n + i] = 0;
        t[1 * n + i] = OP_HASH;                                           // CI
        for (unsigned c = 47; c < 51; c++) t[c * n + i] = consts[C_HASH_INV0];      // State0Inv..State3Inv
        for (unsigned c = 0; c < 16; c++) t[(51 + c) * n + i] = consts[C_RC0 + c]; // Constant0..15 of round 0
      }
    });
  }

  // ---- cascade (cascade.rs:41-66) and lookup (lookup.rs:84-116) tables
  {
    const u64 *cm = aet.cascade_multiplicities, *lm = aet.lookup_multiplicities;
    const size_t cc = aet.cascade_count;
    u64 *t = T + (size_t)COL_CASCADE * n, *u = T + (size_t)COL_LOOKUP * n;
    ex.launch(n, TVM_FILL_BODY(size_t i) {
      if (i < cc) {
        const u64 limb = cm[2 * i], hi = (limb >> 8) & 0xFF, lo = limb & 0xFF;
        t[0 * n + i] = 0; t[1 * n + i] = hi; t[2 * n + i] = lo;
        t[3 * n + i] = consts[C_LUT + hi]; t[4 * n + i] = consts[C_LUT + lo]; t[5 * n + i] = cm[2 * i + 1];
      } else {
        t[0 * n + i] = 1;
        for (unsigned c = 1; c < 6; c++) t[c * n + i] = 0;
      }
      if (i < 256) { u[0 * n + i] = 0; u[1 * n + i] = i; u[2 * n + i] = consts[C_LUT + i]; u[3 * n + i] = lm[i]; }
      else { u[0 * n + i] = 1; u[1 * n + i] = 0; u[2 * n + i] = 0; u[3 * n + i] = 0; }
    });
  }

  // ---- u32 table (u32.rs:105-154, 193-290): one section per entry, in the entries' order
  {
    const size_t cnt = aet.u32_count;
    const u64 *en = aet.u32_entries;
    u64 *t = T + (size_t)COL_U32 * n;
    u64 *total = ex.alloc(1);
    ex.launch(1, TVM_FILL_BODY(size_t) { total[0] = 0; });
    if (cnt) {
      u64 *lens = ex.alloc(cnt), *offs = ex.alloc(cnt);
      ex.launch(cnt, TVM_FILL_BODY(size_t e) { lens[e] = u32_section_len(en[4 * e], en[4 * e + 1], en[4 * e + 2]); });
      ex.exclusive_sum(lens, offs, cnt);
      ex.launch(1, TVM_FILL_BODY(size_t) { total[0] = offs[cnt - 1] + lens[cnt - 1]; });
      info.u32_len = (size_t)ex.read_word(total);
      if (info.u32_len > n) throw std::invalid_argument("the table height is below the u32 table's length");
      ex.launch(cnt, TVM_FILL_BODY(size_t e) {
        const u64 op = en[4 * e], lhs = en[4 * e + 1], rhs = en[4 * e + 2], mult = en[4 * e + 3];
        const size_t base = (size_t)offs[e];
        const unsigned len = (unsigned)lens[e];
        for (unsigned k = 0; k < len; k++) {
          const u64 l = op == OP_POW ? lhs : lhs >> k, r = rhs >> k;
          const size_t row = base + k;
          t[0 * n + row] = k == 0;                       // CopyFlag
          t[1 * n + row] = k;                            // Bits
          t[2 * n + row] = consts[C_BITS33 + (k < 34 ? k : 33)];
          t[3 * n + row] = op;
          t[4 * n + row] = l;
          t[5 * n + row] = inv_or_zero_canon(l);
          t[6 * n + row] = r;
          t[7 * n + row] = inv_or_zero_canon(r);
          t[9 * n + row] = k == 0 ? mult : 0;
        }
        // results from the last row upwards
        u64 res = op == OP_LT ? (len == 1 ? 0 : 2) : (op == OP_LOG2FLOOR ? P - 1 : (op == OP_POW ? 1 : 0));
        t[8 * n + base + len - 1] = res;
        for (unsigned k = len - 1; k-- > 0;) {
          const u64 l = op == OP_POW ? lhs : lhs >> k, r = rhs >> k, lb = l & 1, rb = r & 1, nr = res;
          if (op == OP_LT) {
            if (nr == 2) res = (lb == 0 && rb == 1) ? 1 : ((lb == 1 && rb == 0) ? 0 : (k == 0 ? 0 : 2));
          } else if (op == OP_AND) res = 2 * nr + lb * rb;
          else if (op == OP_LOG2FLOOR) {
            const u64 l_next = lhs >> (k + 1);
            res = l == 0 ? P - 1 : (l_next != 0 ? nr : k);
          } else if (op == OP_POW) {
            res = mul_canon(nr, nr);
            if (rb) res = mul_canon(res, l);
          } else if (op == OP_POP_COUNT) res = nr + lb;
          // split: the result is copied
          t[8 * n + base + k] = res;
        }
      });
    }
    ex.launch(n, TVM_FILL_BODY(size_t i) {
      const size_t tot = (size_t)total[0];
      if (i < tot) return;
      u64 ci = OP_SPLIT, lhs = 0, lhs_inv = 0, res = 0;
      if (tot) {
        ci = t[3 * n + tot - 1]; lhs = t[4 * n + tot - 1]; lhs_inv = t[5 * n + tot - 1]; res = t[8 * n + tot - 1];
        if (ci == OP_LT) res = 2;
      }
      t[0 * n + i] = 0; t[1 * n + i] = 0; t[2 * n + i] = consts[C_BITS33]; t[3 * n + i] = ci; t[4 * n + i] = lhs;
      t[5 * n + i] = lhs_inv; t[6 * n + i] = 0; t[7 * n + i] = 0; t[8 * n + i] = res; t[9 * n + i] = 0;
    });
  }
  return info;
}

}  // namespace fill
}  // namespace tvm

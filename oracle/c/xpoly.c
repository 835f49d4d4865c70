/* CPU oracle, part 2: the array stages of Prover::prove that oracle/stark.py and oracle/stir.py used to run on Python
 * ints (out-of-domain rows, weighted column sums, DEEP, quotient-segment randomisation, dense X-field polynomial
 * arithmetic of STIR).  Same results, plain C + OpenMP over the axes rayon uses in the reference, so that a COMPLETE
 * prove can be timed on the host cores (bench.py --impl reference).
 *
 * TEST INFRASTRUCTURE ONLY (see tvm_oracle.h / oracle/__init__.py).  All values Montgomery form; X-field arrays are
 * interleaved [k][3].  The pure-Python restatements stay in oracle/ (fast=False) and the tests compare the two.
 */
#include "tvm_oracle.h"
#include "field_inline.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MONT_ONE 0xFFFFFFFFULL
static inline u64 fneg(u64 a) { return a ? ORC_P - a : 0; }
static inline xfe xsub(xfe a, xfe b) { xfe r = {fsub(a.c0, b.c0), fsub(a.c1, b.c1), fsub(a.c2, b.c2)}; return r; }
static inline xfe xld(const u64 *p) { xfe r = {p[0], p[1], p[2]}; return r; }
static inline void xst(u64 *p, xfe v) { p[0] = v.c0; p[1] = v.c1; p[2] = v.c2; }
static inline int xis0(xfe a) { return (a.c0 | a.c1 | a.c2) == 0; }
static const xfe X0 = {0, 0, 0};
static const xfe X1 = {MONT_ONE, 0, 0};
static xfe xinv_(xfe a) { u64 i[3] = {a.c0, a.c1, a.c2}, o[3]; orc_xinv(i, o); return xld(o); }
static xfe xpow_(xfe a, u64 e) {
  xfe r = X1;
  while (e) { if (e & 1) r = xmul(r, a); a = xmul(a, a); e >>= 1; }
  return r;
}
static int nthreads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* Montgomery's trick on X-field elements (XFieldElement::batch_inversion, call site master_table.rs:358) */
static void xbatch_inverse(u64 *x, size_t n) {
  if (!n) return;
  u64 *pre = (u64 *)malloc(3 * n * sizeof(u64));
  xfe acc = X1;
  for (size_t i = 0; i < n; i++) { xst(pre + 3 * i, acc); acc = xmul(acc, xld(x + 3 * i)); }
  acc = xinv_(acc);
  for (size_t i = n; i-- > 0;) { xfe t = xmul(acc, xld(pre + 3 * i)); acc = xmul(acc, xld(x + 3 * i)); xst(x + 3 * i, t); }
  free(pre);
}
static void xbatch_inverse_par(u64 *x, size_t n) {
  size_t chunk = 1 << 12;
#pragma omp parallel for schedule(static)
  for (size_t s = 0; s < n; s += chunk) xbatch_inverse(x + 3 * s, (n - s < chunk) ? n - s : chunk);
}

/* ---- dense X-field polynomials (coefficients little-endian) -------------------------------------------------- */
static xfe horner(const u64 *c, size_t len, xfe x) {
  xfe acc = X0;
  for (size_t i = len; i-- > 0;) acc = xadd(xmul(acc, x), xld(c + 3 * i));
  return acc;
}
void orc_xpoly_eval(const u64 *c, size_t len, const u64 x[3], u64 out[3]) {
  xfe xx = xld(x);
  int T = nthreads();
  if (len < 4096 || T == 1) { xst(out, horner(c, len, xx)); return; }
  size_t chunk = (len + T - 1) / T;
  xfe *part = (xfe *)malloc(T * sizeof(xfe));
#pragma omp parallel for schedule(static)
  for (int t = 0; t < T; t++) {
    size_t s = (size_t)t * chunk, e = s + chunk > len ? len : s + chunk;
    part[t] = s < len ? horner(c + 3 * s, e - s, xx) : X0;
  }
  xfe xc = xpow_(xx, chunk), acc = X0;
  for (int t = T; t-- > 0;) acc = xadd(xmul(acc, xc), part[t]);
  free(part);
  xst(out, acc);
}
/* evaluation of a B-field polynomial at an X-field point */
void orc_bpoly_eval_x(const u64 *c, size_t len, const u64 x[3], u64 out[3]) {
  xfe xx = xld(x), acc = X0;
  for (size_t i = len; i-- > 0;) acc = xaddb(xmul(acc, xx), c[i]);
  xst(out, acc);
}
/* stir.rs:1132-1147: out[i] = (c[i*ff .. (i+1)*ff))(r) */
void orc_xpoly_fold(const u64 *c, size_t len, size_t ff, const u64 r[3], u64 *out) {
  xfe rr = xld(r);
  size_t m = (len + ff - 1) / ff;
#pragma omp parallel for schedule(static) if (m >= 1024)
  for (size_t i = 0; i < m; i++) {
    size_t s = i * ff, l = s + ff > len ? len - s : ff;
    xst(out + 3 * i, horner(c + 3 * s, l, rr));
  }
}
void orc_xpoly_mul(const u64 *a, size_t la, const u64 *b, size_t lb, u64 *out) {
  if (!la || !lb) return;
  if (la < lb) { const u64 *t = a; a = b; b = t; size_t tl = la; la = lb; lb = tl; }
  size_t lo = la + lb - 1;
#pragma omp parallel for schedule(static) if (lo >= 1024)
  for (size_t k = 0; k < lo; k++) {
    xfe acc = X0;
    size_t j0 = k >= la ? k - la + 1 : 0, j1 = k < lb - 1 ? k : lb - 1;
    for (size_t j = j0; j <= j1; j++) acc = xadd(acc, xmul(xld(a + 3 * (k - j)), xld(b + 3 * j)));
    xst(out + 3 * k, acc);
  }
}
/* Euclidean division, remainder dropped (the reference's `/`, stir.rs:953-966); num is clobbered; q has ln - ld + 1 words*3 */
void orc_xpoly_div(u64 *num, size_t ln, const u64 *den, size_t ld, u64 *q) {
  if (ln < ld || !ld) return;
  xfe lead_inv = xinv_(xld(den + 3 * (ld - 1)));
  for (size_t k = ln - ld + 1; k-- > 0;) {
    xfe c = xmul(xld(num + 3 * (k + ld - 1)), lead_inv);
    xst(q + 3 * k, c);
    if (!xis0(c))
      for (size_t j = 0; j < ld; j++) xst(num + 3 * (k + j), xsub(xld(num + 3 * (k + j)), xmul(c, xld(den + 3 * j))));
  }
}
/* Polynomial::zerofier: prod (X - p_i); out has k + 1 coefficients */
void orc_xzerofier(const u64 *pts, size_t k, u64 *out) {
  xfe *z = (xfe *)malloc((k + 1) * sizeof(xfe));
  z[0] = X1;
  for (size_t i = 0; i < k; i++) {
    xfe p = xld(pts + 3 * i);
    z[i + 1] = z[i];
    for (size_t j = i; j > 0; j--) z[j] = xsub(z[j - 1], xmul(z[j], p));
    z[0] = xsub(X0, xmul(z[0], p));
  }
  for (size_t j = 0; j <= k; j++) xst(out + 3 * j, z[j]);
  free(z);
}
/* Polynomial::interpolate (Lagrange), k points; out has k coefficients */
void orc_xinterpolate(const u64 *xs, const u64 *ys, size_t k, u64 *out) {
  u64 *z = (u64 *)malloc(3 * (k + 1) * sizeof(u64));
  orc_xzerofier(xs, k, z);
  int T = nthreads();
  u64 *acc = (u64 *)calloc((size_t)T * 3 * k, sizeof(u64));
#pragma omp parallel
  {
#ifdef _OPENMP
    int t = omp_get_thread_num();
#else
    int t = 0;
#endif
    u64 *mine = acc + (size_t)t * 3 * k;
    xfe *q = (xfe *)malloc(k * sizeof(xfe));
#pragma omp for schedule(static)
    for (size_t i = 0; i < k; i++) {
      xfe xi = xld(xs + 3 * i), a = X0;
      for (size_t j = k; j >= 1; j--) { a = xadd(xld(z + 3 * j), xmul(a, xi)); q[j - 1] = a; }   /* z / (X - xi) */
      xfe denom = X0;
      for (size_t j = k; j-- > 0;) denom = xadd(xmul(denom, xi), q[j]);
      xfe s = xmul(xld(ys + 3 * i), xinv_(denom));
      for (size_t j = 0; j < k; j++) xst(mine + 3 * j, xadd(xld(mine + 3 * j), xmul(s, q[j])));
    }
    free(q);
  }
  for (size_t j = 0; j < k; j++) {
    xfe s = X0;
    for (int t = 0; t < T; t++) s = xadd(s, xld(acc + (size_t)t * 3 * k + 3 * j));
    xst(out + 3 * j, s);
  }
  free(acc); free(z);
}
/* dst[j] += w * src[j] */
void orc_xpoly_axpy(u64 *dst, const u64 *src, size_t len, const u64 w[3]) {
  xfe ww = xld(w);
#pragma omp parallel for schedule(static) if (len >= 4096)
  for (size_t j = 0; j < len; j++) xst(dst + 3 * j, xadd(xld(dst + 3 * j), xmul(ww, xld(src + 3 * j))));
}
/* dst[j] += src[j] * scale * arg^j   (B-field scale, arg): f(X) -> scale * f(arg X), stark.rs:1302-1356 */
void orc_xpoly_add_scaled_arg(u64 *dst, const u64 *src, size_t len, u64 scale, u64 arg) {
  size_t chunk = 1 << 12;
#pragma omp parallel for schedule(static) if (len >= 2 * chunk)
  for (size_t s = 0; s < len; s += chunk) {
    u64 f = fmul(scale, orc_pow(arg, s));
    size_t e = s + chunk > len ? len : s + chunk;
    for (size_t j = s; j < e; j++) { xst(dst + 3 * j, xadd(xld(dst + 3 * j), xmulb(xld(src + 3 * j), f))); f = fmul(f, arg); }
  }
}

/* ---- master-table stages ------------------------------------------------------------------------------------ */
/* randomized_column_interpolant of every column (master_table.rs:392-403): coefficients [ncols][2n] */
void orc_interpolants_table(const u64 *trace, unsigned log2n, size_t ncols, const u64 *rand, size_t h, u64 *out) {
  size_t n = (size_t)1 << log2n;
#pragma omp parallel for schedule(dynamic)
  for (size_t c = 0; c < ncols; c++) {
    u64 *o = out + c * 2 * n;
    memcpy(o, trace + c * n, n * sizeof(u64));
    orc_intt(o, log2n);
    memset(o + n, 0, n * sizeof(u64));
    for (size_t i = 0; rand && i < h; i++) { u64 r = rand[c * h + i]; o[i] = fsub(o[i], r); o[n + i] = fadd(o[n + i], r); }
  }
}
/* barycentric weights of the trace domain (offset 1) for the point alpha, master_table.rs:348-390:
 * dods[i] = d_i / (alpha - d_i), denom_inv = 1 / sum_i dods[i] */
void orc_bary_weights(unsigned log2n, const u64 alpha[3], u64 *dods, u64 denom_inv[3]) {
  size_t n = (size_t)1 << log2n, chunk = 1 << 12;
  u64 g = orc_root_of_unity(log2n);
  xfe a = xld(alpha);
#pragma omp parallel for schedule(static)
  for (size_t s = 0; s < n; s += chunk) {
    u64 d = orc_pow(g, s);
    size_t e = s + chunk > n ? n : s + chunk;
    for (size_t i = s; i < e; i++) { xfe t = {fsub(a.c0, d), a.c1, a.c2}; xst(dods + 3 * i, t); d = fmul(d, g); }
  }
  xbatch_inverse_par(dods, n);
  int T = nthreads();
  xfe *part = (xfe *)calloc(T, sizeof(xfe));
#pragma omp parallel
  {
#ifdef _OPENMP
    int t = omp_get_thread_num();
#else
    int t = 0;
#endif
    xfe acc = X0;
#pragma omp for schedule(static)
    for (size_t s = 0; s < n; s += chunk) {
      u64 d = orc_pow(g, s);
      size_t e = s + chunk > n ? n : s + chunk;
      for (size_t i = s; i < e; i++) { xfe v = xmulb(xld(dods + 3 * i), d); xst(dods + 3 * i, v); acc = xadd(acc, v); d = fmul(d, g); }
    }
    part[t] = acc;
  }
  xfe den = X0;
  for (int t = 0; t < T; t++) den = xadd(den, part[t]);
  free(part);
  xst(denom_inv, xinv_(den));
}
/* out[c] = (sum_i col_c[i] dods[i]) denom_inv + (alpha^n - 1) rand_c(alpha);  xf = 1: B-field columns [ncols][n], rand
 * [ncols][h];  xf = 3: X-field columns as planes [3 ncols][n], rand planes [3 ncols][h] */
void orc_ood_row(const u64 *cols, size_t ncols, unsigned log2n, int xf, const u64 *dods, const u64 denom_inv[3], const u64 *rand,
                 size_t h, const u64 alpha[3], u64 *out) {
  size_t n = (size_t)1 << log2n;
  xfe a = xld(alpha), di = xld(denom_inv);
  xfe zerofier = xsub(xpow_(a, n), X1);
#pragma omp parallel for schedule(dynamic)
  for (size_t c = 0; c < ncols; c++) {
    xfe num = X0, r = X0;
    if (xf == 1) {
      const u64 *col = cols + c * n;
      for (size_t i = 0; i < n; i++) num = xadd(num, xmulb(xld(dods + 3 * i), col[i]));
      for (size_t i = h; i-- > 0;) r = xaddb(xmul(r, a), rand[c * h + i]);
    } else {
      const u64 *p0 = cols + (3 * c) * n, *p1 = p0 + n, *p2 = p1 + n;
      for (size_t i = 0; i < n; i++) { xfe v = {p0[i], p1[i], p2[i]}; num = xadd(num, xmul(v, xld(dods + 3 * i))); }
      const u64 *r0 = rand + (3 * c) * h, *r1 = r0 + h, *r2 = r1 + h;
      for (size_t i = h; i-- > 0;) { xfe v = {r0[i], r1[i], r2[i]}; r = xadd(xmul(r, a), v); }
    }
    xst(out + 3 * c, xadd(xmul(num, di), xmul(zerofier, r)));
  }
}
/* weighted_sum_of_columns on the column interpolants (master_table.rs:512-542): out[j] = sum_c w_c coef_c[j];
 * main_coef [nmain][len] B-field, aux_coef [3 naux][len] planes, w [(nmain + naux)][3] */
void orc_weighted_colsum(const u64 *main_coef, size_t nmain, const u64 *aux_coef, size_t naux, size_t len, const u64 *w, u64 *out) {
  size_t blk = 1 << 10;
#pragma omp parallel for schedule(static)
  for (size_t s = 0; s < len; s += blk) {
    size_t e = s + blk > len ? len : s + blk;
    for (size_t j = s; j < e; j++) xst(out + 3 * j, X0);
    for (size_t c = 0; c < nmain; c++) {
      xfe wc = xld(w + 3 * c);
      const u64 *col = main_coef + c * len;
      for (size_t j = s; j < e; j++)
        if (col[j]) xst(out + 3 * j, xadd(xld(out + 3 * j), xmulb(wc, col[j])));
    }
    for (size_t q = 0; q < naux; q++) {
      xfe wc = xld(w + 3 * (nmain + q));
      const u64 *p0 = aux_coef + (3 * q) * len, *p1 = p0 + len, *p2 = p1 + len;
      for (size_t j = s; j < e; j++) {
        xfe v = {p0[j], p1[j], p2[j]};
        if (!xis0(v)) xst(out + 3 * j, xadd(xld(out + 3 * j), xmul(wc, v)));
      }
    }
  }
}
/* DEEP (stark.rs:1360-1379, 2096-2103) of the 4 codewords and their weighted sum (612-626):
 * out[i] = sum_k w_k (cw_k[i] - value_k) / (x_i - point_k),  x_i = offset g^i;  cw[0] is used for k = 0 and 1 */
void orc_deep_combination(const u64 *cw_ma, const u64 *cw_p, const u64 *cw_r, unsigned log2N, u64 offset, const u64 *points,
                          const u64 *values, const u64 *weights, u64 *out) {
  size_t N = (size_t)1 << log2N, chunk = 1 << 11;
  u64 g = orc_root_of_unity(log2N);
  const u64 *cws[4] = {cw_ma, cw_ma, cw_p, cw_r};
#pragma omp parallel
  {
    u64 *inv = (u64 *)malloc(4 * 3 * chunk * sizeof(u64));
#pragma omp for schedule(static)
    for (size_t s = 0; s < N; s += chunk) {
      size_t e = s + chunk > N ? N : s + chunk, m = e - s;
      u64 x = fmul(offset, orc_pow(g, s));
      for (size_t i = 0; i < m; i++) {
        for (int k = 0; k < 4; k++) {
          xfe p = xld(points + 3 * k);
          xfe t = {fsub(x, p.c0), fneg(p.c1), fneg(p.c2)};
          xst(inv + 3 * (k * m + i), t);
        }
        x = fmul(x, g);
      }
      xbatch_inverse(inv, 4 * m);
      for (size_t i = 0; i < m; i++) {
        xfe acc = X0;
        for (int k = 0; k < 4; k++) {
          xfe d = xmul(xsub(xld(cws[k] + 3 * (s + i)), xld(values + 3 * k)), xld(inv + 3 * (k * m + i)));
          acc = xadd(acc, xmul(d, xld(weights + 3 * k)));
        }
        xst(out + 3 * (s + i), acc);
      }
    }
    free(inv);
  }
}

/* CPU oracle (plain C + OpenMP) — field, NTT, LDE, Tip5, Merkle.
 *
 * TEST INFRASTRUCTURE ONLY (see tvm_oracle.h / oracle/__init__.py).
 *
 * Restates the algorithms of twenty-first 2.0.0 (un-vendored dependency of the
 * reference, Cargo.toml:104) at the call sites listed in SURVEY.md §8(c), and the
 * reference's own arithmetic_domain.rs:141-189, master_table.rs:258-322,392-503.
 * Parallel axes are the ones rayon uses in the reference: columns for LDE
 * (master_table.rs:280-314), rows for hashing (458-464), tree levels for Merkle.
 */
#include "tvm_oracle.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
#define P 0xFFFFFFFF00000001ULL
#define MONT_ONE 0xFFFFFFFFULL            /* R mod p */
#define MONT_R2 0xFFFFFFFE00000001ULL     /* R^2 mod p */
#define ROOT_2_32_CANON 1753635133440165772ULL

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* Montgomery reduction x*R^-1 mod p, x < p*2^64 (Pornin's formulation for this prime). */
static inline u64 montyred(u128 x) {
  u64 xl = (u64)x, xh = (u64)(x >> 64);
  u64 a; unsigned e = __builtin_add_overflow(xl, xl << 32, &a);
  u64 b = a - (a >> 32) - e;
  u64 r; unsigned c = __builtin_sub_overflow(xh, b, &r);
  return r - (0xFFFFFFFFULL * c);
}
static inline u64 fmul(u64 a, u64 b) { return montyred((u128)a * b); }
static inline u64 fadd(u64 a, u64 b) {
  u64 s; unsigned c = __builtin_add_overflow(a, b, &s);
  if (c || s >= P) s -= P;
  return s;
}
static inline u64 fsub(u64 a, u64 b) {
  u64 d; unsigned c = __builtin_sub_overflow(a, b, &d);
  if (c) d += P;
  return d;
}
static inline u64 fneg(u64 a) { return a ? P - a : 0; }
static inline u64 to_mont1(u64 x) { return montyred((u128)x * MONT_R2); }
static inline u64 from_mont1(u64 x) { return montyred((u128)x); }

void orc_to_mont(u64 *x, size_t n) { for (size_t i = 0; i < n; i++) x[i] = to_mont1(x[i] % P); }
void orc_from_mont(u64 *x, size_t n) { for (size_t i = 0; i < n; i++) x[i] = from_mont1(x[i]); }
u64 orc_mul(u64 a, u64 b) { return fmul(a, b); }
u64 orc_pow(u64 a, u64 e) {
  u64 r = MONT_ONE;
  while (e) { if (e & 1) r = fmul(r, a); a = fmul(a, a); e >>= 1; }
  return r;
}
u64 orc_inv(u64 a) { return orc_pow(a, P - 2); }
u64 orc_root_of_unity(unsigned log2n) {
  u64 r = to_mont1(ROOT_2_32_CANON);
  for (unsigned i = log2n; i < 32; i++) r = fmul(r, r);
  return r;
}

/* ---- X-field, X^3 = X - 1 (specification/src/isa.md:8) ---- */
void orc_xmul(const u64 a[3], const u64 b[3], u64 out[3]) {
  u64 d0 = fmul(a[0], b[0]);
  u64 d1 = fadd(fmul(a[0], b[1]), fmul(a[1], b[0]));
  u64 d2 = fadd(fadd(fmul(a[0], b[2]), fmul(a[1], b[1])), fmul(a[2], b[0]));
  u64 d3 = fadd(fmul(a[1], b[2]), fmul(a[2], b[1]));
  u64 d4 = fmul(a[2], b[2]);
  out[0] = fsub(d0, d3);
  out[1] = fsub(fadd(d1, d3), d4);
  out[2] = fadd(d2, d4);
}
void orc_xinv(const u64 a[3], u64 out[3]) {
  /* a^-1 = a^(p^3-2); use Frobenius-free square-and-multiply over 192 bits:
   * p^3 - 2 as a 192-bit exponent, little-endian words. */
  /* p^3 = (2^64 - 2^32 + 1)^3 ; computed once */
  static u64 e[3]; static int init = 0;
  if (!init) {
    u128 p = P;
    /* p^2 */
    u128 lo = (u128)(u64)p * (u64)p;          /* p fits in 64 bits */
    u64 p2_0 = (u64)lo, p2_1 = (u64)(lo >> 64);
    /* p^3 = p2 * p */
    u128 t0 = (u128)p2_0 * P;
    u128 t1 = (u128)p2_1 * P + (u64)(t0 >> 64);
    u64 w0 = (u64)t0, w1 = (u64)t1, w2 = (u64)(t1 >> 64);
    /* minus 2 */
    u64 b = w0 < 2; w0 -= 2;
    u64 b2 = w1 < b; w1 -= b; w2 -= b2;
    e[0] = w0; e[1] = w1; e[2] = w2; init = 1;
  }
  u64 r[3] = {MONT_ONE, 0, 0}, base[3] = {a[0], a[1], a[2]}, t[3];
  for (int w = 0; w < 3; w++)
    for (int bit = 0; bit < 64; bit++) {
      if ((e[w] >> bit) & 1) { orc_xmul(r, base, t); memcpy(r, t, 24); }
      orc_xmul(base, base, t); memcpy(base, t, 24);
    }
  memcpy(out, r, 24);
}

/* ---- NTT ---- */
static void bitrev_permute(u64 *x, unsigned log2n, size_t stride) {
  size_t n = (size_t)1 << log2n;
  for (size_t i = 0; i < n; i++) {
    size_t j = 0;
    for (unsigned b = 0; b < log2n; b++) j |= ((i >> b) & 1) << (log2n - 1 - b);
    if (j > i) { u64 t = x[i * stride]; x[i * stride] = x[j * stride]; x[j * stride] = t; }
  }
}
static void ntt_core(u64 *x, unsigned log2n, u64 omega, size_t stride) {
  size_t n = (size_t)1 << log2n;
  if (n == 1) return;
  bitrev_permute(x, log2n, stride);
  u64 *tw = (u64 *)malloc((n / 2) * sizeof(u64));
  tw[0] = MONT_ONE;
  for (size_t i = 1; i < n / 2; i++) tw[i] = fmul(tw[i - 1], omega);
  for (unsigned s = 1; s <= log2n; s++) {
    size_t m = (size_t)1 << s, half = m >> 1, step = n / m;
    for (size_t k = 0; k < n; k += m)
      for (size_t j = 0; j < half; j++) {
        u64 u = x[(k + j) * stride];
        u64 v = fmul(x[(k + j + half) * stride], tw[j * step]);
        x[(k + j) * stride] = fadd(u, v);
        x[(k + j + half) * stride] = fsub(u, v);
      }
  }
  free(tw);
}
void orc_ntt(u64 *x, unsigned log2n) { ntt_core(x, log2n, orc_root_of_unity(log2n), 1); }
void orc_intt(u64 *x, unsigned log2n) {
  size_t n = (size_t)1 << log2n;
  ntt_core(x, log2n, orc_inv(orc_root_of_unity(log2n)), 1);
  u64 ninv = orc_inv(to_mont1((u64)n));
  for (size_t i = 0; i < n; i++) x[i] = fmul(x[i], ninv);
}
void orc_xntt(u64 *x, unsigned log2n) {
  u64 w = orc_root_of_unity(log2n);
  for (int c = 0; c < 3; c++) ntt_core(x + c, log2n, w, 3);
}
void orc_xintt(u64 *x, unsigned log2n) {
  size_t n = (size_t)1 << log2n;
  u64 w = orc_inv(orc_root_of_unity(log2n));
  for (int c = 0; c < 3; c++) ntt_core(x + c, log2n, w, 3);
  u64 ninv = orc_inv(to_mont1((u64)n));
  for (size_t i = 0; i < 3 * n; i++) x[i] = fmul(x[i], ninv);
}

/* arithmetic_domain.rs:141-170 — polynomials longer than the domain are folded
 * chunk-wise with offset^(chunk*len) scaling. */
void orc_coset_evaluate(const u64 *coef, size_t ncoef, u64 offset, unsigned log2n, u64 *out) {
  size_t n = (size_t)1 << log2n;
  u64 *tmp = (u64 *)malloc(n * sizeof(u64));
  memset(out, 0, n * sizeof(u64));
  size_t nchunks = (ncoef + n - 1) / n;
  for (size_t ch = 0; ch < nchunks; ch++) {
    size_t len = ncoef - ch * n < n ? ncoef - ch * n : n;
    u64 acc = MONT_ONE;
    for (size_t i = 0; i < len; i++) { tmp[i] = fmul(coef[ch * n + i], acc); acc = fmul(acc, offset); }
    for (size_t i = len; i < n; i++) tmp[i] = 0;
    orc_ntt(tmp, log2n);
    if (ch == 0) memcpy(out, tmp, n * sizeof(u64));
    else {
      u64 so = orc_pow(offset, (u64)ch * n);
      for (size_t i = 0; i < n; i++) out[i] = fadd(out[i], fmul(tmp[i], so));
    }
  }
  free(tmp);
}
void orc_coset_interpolate(const u64 *vals, u64 offset, unsigned log2n, u64 *coef) {
  size_t n = (size_t)1 << log2n;
  memcpy(coef, vals, n * sizeof(u64));
  orc_intt(coef, log2n);
  u64 oi = orc_inv(offset), acc = MONT_ONE;
  for (size_t i = 0; i < n; i++) { coef[i] = fmul(coef[i], acc); acc = fmul(acc, oi); }
}

/* master_table.rs:392-403 (randomized_column_interpolant; trace-domain offset is 1 so
 * zerofier = X^n - 1, arithmetic_domain.rs:250-269) then 309-314 (evaluate on the
 * evaluation domain). */
void orc_lde_column(const u64 *trace, unsigned log2_trace, const u64 *randomizer, size_t num_rand,
                    u64 eval_offset, unsigned log2_eval, u64 *out, u64 *coef_out) {
  size_t n = (size_t)1 << log2_trace;
  u64 *coef = (u64 *)calloc(2 * n, sizeof(u64));
  orc_coset_interpolate(trace, MONT_ONE, log2_trace, coef);
  for (size_t i = 0; i < num_rand; i++) {
    coef[n + i] = fadd(coef[n + i], randomizer[i]);
    coef[i] = fsub(coef[i], randomizer[i]);
  }
  orc_coset_evaluate(coef, n + num_rand, eval_offset, log2_eval, out);
  if (coef_out) memcpy(coef_out, coef, 2 * n * sizeof(u64));
  free(coef);
}

/* ---- Tip5 (tips/tip-0005/tip-0005.md:21-81) ---- */
static const uint16_t MDS_FIRST_COLUMN[16] = {61402, 1108, 28750, 33823, 7454, 43244, 53865, 12034,
                                              56951, 27521, 41351, 40901, 12021, 59689, 26798, 17845};
static uint8_t LOOKUP[256];
static u64 ROUND_CONSTANTS[80]; /* Montgomery form; filled by orc_tip5_set_round_constants */
static int tip5_ready = 0;

/* The round constants are BLAKE3-derived (tip-0005.md:72); the Python side derives
 * them with the blake3 module and hands them over (canonical form). */
void orc_tip5_set_round_constants(const u64 canon[80]) {
  for (int i = 0; i < 80; i++) ROUND_CONSTANTS[i] = to_mont1(canon[i]);
  for (int x = 0; x < 256; x++) {
    unsigned y = (unsigned)(x + 1) % 257;
    unsigned c = (unsigned)(((unsigned long)y * y % 257) * y % 257);
    LOOKUP[x] = (uint8_t)((c + 256) % 257);
  }
  tip5_ready = 1;
}

static inline u64 reduce_small_hi(u128 x) { /* x < 2^96: plain x mod p */
  u64 lo = (u64)x, hi = (u64)(x >> 64); /* hi < 2^32 */
  u64 t = hi * 0xFFFFFFFFULL;           /* hi * (2^32 - 1) < 2^64 */
  u64 s; unsigned c = __builtin_add_overflow(lo, t, &s);
  if (c) s += 0xFFFFFFFFULL;            /* 2^64 = 2^32 - 1 mod p; cannot overflow again */
  if (s >= P) s -= P;
  return s;
}

void orc_tip5_permutation(u64 s[16]) {
  if (!tip5_ready) abort();
  for (int rnd = 0; rnd < 5; rnd++) {
    for (int i = 0; i < 4; i++) {
      u64 v = s[i], o = 0;
      for (int b = 0; b < 8; b++) o |= (u64)LOOKUP[(v >> (8 * b)) & 0xFF] << (8 * b);
      s[i] = o;
    }
    for (int i = 4; i < 16; i++) {
      u64 x = s[i], x2 = fmul(x, x), x3 = fmul(x2, x), x4 = fmul(x2, x2);
      s[i] = fmul(x3, x4);
    }
    u64 t[16];
    for (int r = 0; r < 16; r++) {
      u128 acc = 0;
      for (int c = 0; c < 16; c++) acc += (u128)MDS_FIRST_COLUMN[(r - c) & 15] * s[c];
      t[r] = reduce_small_hi(acc);
    }
    for (int i = 0; i < 16; i++) s[i] = fadd(t[i], ROUND_CONSTANTS[16 * rnd + i]);
  }
}

void orc_hash_varlen(const u64 *w, size_t n, u64 digest[5]) {
  u64 s[16] = {0};
  size_t full = n / 10;
  for (size_t i = 0; i < full; i++) { memcpy(s, w + 10 * i, 80); orc_tip5_permutation(s); }
  size_t rem = n - 10 * full;
  for (size_t i = 0; i < rem; i++) s[i] = w[10 * full + i];
  s[rem] = MONT_ONE;
  for (size_t i = rem + 1; i < 10; i++) s[i] = 0;
  orc_tip5_permutation(s);
  memcpy(digest, s, 40);
}
void orc_hash_pair(const u64 l[5], const u64 r[5], u64 digest[5]) {
  u64 s[16];
  memcpy(s, l, 40); memcpy(s + 5, r, 40);
  for (int i = 10; i < 16; i++) s[i] = MONT_ONE;
  orc_tip5_permutation(s);
  memcpy(digest, s, 40);
}

void orc_hash_rows_colmajor(const u64 *table, size_t nrows, size_t ncols, size_t col_stride, u64 *digests) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < nrows; i++) {
    u64 s[16] = {0};
    size_t c = 0;
    while (c + 10 <= ncols) {
      for (int k = 0; k < 10; k++) s[k] = table[(c + k) * col_stride + i];
      orc_tip5_permutation(s);
      c += 10;
    }
    size_t rem = ncols - c;
    for (size_t k = 0; k < rem; k++) s[k] = table[(c + k) * col_stride + i];
    s[rem] = MONT_ONE;
    for (size_t k = rem + 1; k < 10; k++) s[k] = 0;
    orc_tip5_permutation(s);
    memcpy(digests + 5 * i, s, 40);
  }
}

void orc_merkle_build(const u64 *leaves, size_t nleaves, u64 *nodes) {
  memcpy(nodes + 5 * nleaves, leaves, nleaves * 40);
  memset(nodes, 0, 40);
  for (size_t lvl = nleaves / 2; lvl >= 1; lvl /= 2) {
#pragma omp parallel for schedule(static) if (lvl >= 256)
    for (size_t i = lvl; i < 2 * lvl; i++) orc_hash_pair(nodes + 5 * (2 * i), nodes + 5 * (2 * i + 1), nodes + 5 * i);
  }
}

/* batched LDE over columns (rayon axis of master_table.rs:280-314) */
void orc_lde_table(const u64 *trace_colmajor, unsigned log2_trace, size_t ncols, const u64 *randomizers,
                   size_t num_rand, u64 eval_offset, unsigned log2_eval, u64 *out_colmajor) {
  size_t n = (size_t)1 << log2_trace, m = (size_t)1 << log2_eval;
#pragma omp parallel for schedule(dynamic)
  for (size_t c = 0; c < ncols; c++)
    orc_lde_column(trace_colmajor + c * n, log2_trace, randomizers ? randomizers + c * num_rand : NULL,
                   randomizers ? num_rand : 0, eval_offset, log2_eval, out_colmajor + c * m, NULL);
}

/* ---- AIR quotient (master_table.rs:1194-1363) ------------------------------------------- */
void orc_air_eval_init(const u64 *mc, const u64 *ac, const u64 *mn, const u64 *an, const u64 *ch, u64 *out);
void orc_air_eval_cons(const u64 *mc, const u64 *ac, const u64 *mn, const u64 *an, const u64 *ch, u64 *out);
void orc_air_eval_tran(const u64 *mc, const u64 *ac, const u64 *mn, const u64 *an, const u64 *ch, u64 *out);
void orc_air_eval_term(const u64 *mc, const u64 *ac, const u64 *mn, const u64 *an, const u64 *ch, u64 *out);
extern const int ORC_AIR_NUM_CONSTRAINTS[4];

static void batch_inverse(u64 *x, size_t n) {
  if (!n) return;
  u64 *pre = (u64 *)malloc(n * sizeof(u64));
  u64 acc = MONT_ONE;
  for (size_t i = 0; i < n; i++) { pre[i] = acc; acc = fmul(acc, x[i]); }
  acc = orc_inv(acc);
  for (size_t i = n; i-- > 0;) { u64 t = fmul(acc, pre[i]); acc = fmul(acc, x[i]); x[i] = t; }
  free(pre);
}
void orc_batch_inverse(u64 *x, size_t n) { batch_inverse(x, n); }

/* main: column-major [nmain][N] in natural domain order; aux: [3*naux][N] (X-field columns planar).
 * Row i's successor is row (i + unit_distance) mod N.  weights: 604 X-field; out: [N][3]. */
void orc_air_quotient(const u64 *main, size_t nmain, const u64 *aux, size_t naux3, size_t N, unsigned log2_trace,
                      u64 domain_offset, const u64 *ch, const u64 *w, u64 *out) {
  unsigned log2N = 0; while (((size_t)1 << log2N) < N) log2N++;
  size_t n = (size_t)1 << log2_trace, unit = N / n;
  u64 g = orc_root_of_unity(log2N), wn_inv = orc_inv(orc_root_of_unity(log2_trace));
  u64 *zi = (u64 *)malloc(4 * N * sizeof(u64));
  u64 *z_init = zi, *z_cons = zi + N, *z_tran = zi + 2 * N, *z_term = zi + 3 * N;
  /* x_i^n = offset^n (g^n)^i takes only N/n values; the three inversions are batched per chunk (Montgomery's trick) */
  const size_t zchunk = 1 << 12;
  u64 off_n = orc_pow(domain_offset, n), g_n = orc_pow(g, n);
#pragma omp parallel for schedule(static)
  for (size_t s = 0; s < N; s += zchunk) {
    size_t e = s + zchunk > N ? N : s + zchunk;
    u64 x = fmul(domain_offset, orc_pow(g, s)), xn = fmul(off_n, orc_pow(g_n, s % unit));
    for (size_t i = s; i < e; i++) {
      z_init[i] = fsub(x, MONT_ONE);
      z_cons[i] = fsub(xn, MONT_ONE);
      z_term[i] = fsub(x, wn_inv);
      z_tran[i] = z_term[i];
      x = fmul(x, g); xn = fmul(xn, g_n);
    }
    batch_inverse(z_init + s, e - s); batch_inverse(z_cons + s, e - s); batch_inverse(z_term + s, e - s);
    for (size_t i = s; i < e; i++) z_tran[i] = fmul(z_tran[i], z_cons[i]);
  }
  const int nc[4] = {ORC_AIR_NUM_CONSTRAINTS[0], ORC_AIR_NUM_CONSTRAINTS[1], ORC_AIR_NUM_CONSTRAINTS[2], ORC_AIR_NUM_CONSTRAINTS[3]};
#pragma omp parallel
  {
    u64 *mc = (u64 *)malloc((2 * nmain + 2 * naux3) * sizeof(u64));
    u64 *mn = mc + nmain, *ac = mn + nmain, *an = ac + naux3;
    u64 *vals = (u64 *)malloc(3 * 512 * sizeof(u64));
#pragma omp for schedule(static)
    for (size_t i = 0; i < N; i++) {
      size_t j = (i + unit) % N;
      for (size_t c = 0; c < nmain; c++) { mc[c] = main[c * N + i]; mn[c] = main[c * N + j]; }
      for (size_t c = 0; c < naux3; c++) { ac[c] = aux[c * N + i]; an[c] = aux[c * N + j]; }
      u64 acc[3] = {0, 0, 0};
      size_t off = 0;
      for (int cat = 0; cat < 4; cat++) {
        if (cat == 0) orc_air_eval_init(mc, ac, mn, an, ch, vals);
        else if (cat == 1) orc_air_eval_cons(mc, ac, mn, an, ch, vals);
        else if (cat == 2) orc_air_eval_tran(mc, ac, mn, an, ch, vals);
        else orc_air_eval_term(mc, ac, mn, an, ch, vals);
        u64 s[3] = {0, 0, 0}, t[3];
        for (int k = 0; k < nc[cat]; k++) {
          orc_xmul(w + 3 * (off + k), vals + 3 * k, t);
          s[0] = fadd(s[0], t[0]); s[1] = fadd(s[1], t[1]); s[2] = fadd(s[2], t[2]);
        }
        off += nc[cat];
        u64 z = zi[cat * N + i];
        acc[0] = fadd(acc[0], fmul(s[0], z)); acc[1] = fadd(acc[1], fmul(s[1], z)); acc[2] = fadd(acc[2], fmul(s[2], z));
      }
      out[3 * i] = acc[0]; out[3 * i + 1] = acc[1]; out[3 * i + 2] = acc[2];
    }
    free(mc); free(vals);
  }
  free(zi);
}

/* row evaluation of one category (for tests): out has 3 words per constraint */
void orc_air_eval_category(int cat, const u64 *mc, const u64 *ac, const u64 *mn, const u64 *an, const u64 *ch, u64 *out) {
  if (cat == 0) orc_air_eval_init(mc, ac, mn, an, ch, out);
  else if (cat == 1) orc_air_eval_cons(mc, ac, mn, an, ch, out);
  else if (cat == 2) orc_air_eval_tran(mc, ac, mn, an, ch, out);
  else orc_air_eval_term(mc, ac, mn, an, ch, out);
}

/* ---- X-field vector helpers (elements interleaved [n][3]) -------------------------------- */
/* FRI fold: fri.rs:349-366 */
void orc_fri_fold(const u64 *cw, size_t n, u64 domain_offset, const u64 chal[3], u64 *out) {
  unsigned log2n = 0; while (((size_t)1 << log2n) < n) log2n++;
  u64 g = orc_root_of_unity(log2n);
  u64 *inv = (u64 *)malloc((n / 2) * sizeof(u64));
  u64 x = domain_offset;
  for (size_t i = 0; i < n / 2; i++) { inv[i] = x; x = fmul(x, g); }
  batch_inverse(inv, n / 2);
  u64 two_inv = orc_inv(fadd(MONT_ONE, MONT_ONE));
#pragma omp parallel for schedule(static) if (n >= 4096)
  for (size_t i = 0; i < n / 2; i++) {
    u64 s[3] = {fmul(chal[0], inv[i]), fmul(chal[1], inv[i]), fmul(chal[2], inv[i])};
    u64 one_plus[3] = {fadd(MONT_ONE, s[0]), s[1], s[2]};
    u64 one_minus[3] = {fsub(MONT_ONE, s[0]), fneg(s[1]), fneg(s[2])};
    u64 l[3], r[3];
    orc_xmul(one_plus, cw + 3 * i, l);
    orc_xmul(one_minus, cw + 3 * (n / 2 + i), r);
    for (int d = 0; d < 3; d++) out[3 * i + d] = fmul(fadd(l[d], r[d]), two_inv);
  }
  free(inv);
}

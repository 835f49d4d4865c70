/* CPU oracle (plain C) for the Triton VM Stark::prove() hot path.
 *
 * TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.  Only tests/, smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * All arrays hold field elements in the MONTGOMERY representation (R = 2^64),
 * the in-memory form of twenty-first's BFieldElement (reference
 * triton-constraint-builder/src/codegen.rs:926-932), unless a function name says
 * "canon".  X-field elements are 3 consecutive words (c0,c1,c2), digests 5 words.
 */
#ifndef TVM_ORACLE_H
#define TVM_ORACLE_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef uint64_t u64;

/* field */
void orc_to_mont(u64 *x, size_t n);
void orc_from_mont(u64 *x, size_t n);
u64  orc_mul(u64 a, u64 b);            /* Montgomery product */
u64  orc_pow(u64 a, u64 e);
u64  orc_inv(u64 a);
u64  orc_root_of_unity(unsigned log2n); /* Montgomery form */
void orc_xmul(const u64 a[3], const u64 b[3], u64 out[3]);
void orc_xinv(const u64 a[3], u64 out[3]);

/* NTT: natural order in / natural order out, ntt(x)[k] = sum_j x_j w^{jk} */
void orc_ntt(u64 *x, unsigned log2n);
void orc_intt(u64 *x, unsigned log2n);
/* strided X-field variants: element i is x[3*i + {0,1,2}] */
void orc_xntt(u64 *x, unsigned log2n);
void orc_xintt(u64 *x, unsigned log2n);
/* coset evaluate: coeffs (ncoef <= n after folding) -> values on offset*<w_n> */
void orc_coset_evaluate(const u64 *coef, size_t ncoef, u64 offset, unsigned log2n, u64 *out);
void orc_coset_interpolate(const u64 *vals, u64 offset, unsigned log2n, u64 *coef_out);

/* LDE of one trace column: interpolate on trace domain (offset 1), add
 * zerofier*randomizer, evaluate on offset*<w_{2^log2_eval}>.
 * master_table.rs:258-322, 392-434 */
void orc_lde_column(const u64 *trace, unsigned log2_trace, const u64 *randomizer, size_t num_rand,
                    u64 eval_offset, unsigned log2_eval, u64 *out, u64 *coef_out /* 2*trace_len or NULL */);

/* Tip5 */
void orc_tip5_permutation(u64 state[16]);
void orc_hash_varlen(const u64 *words, size_t n, u64 digest[5]);
void orc_hash_pair(const u64 l[5], const u64 r[5], u64 digest[5]);
/* table is column-major [ncols][nrows] (stride between columns = col_stride words);
 * row i = (table[c*col_stride + i])_c ; digests out [nrows][5]  (master_table.rs:455-465) */
void orc_hash_rows_colmajor(const u64 *table, size_t nrows, size_t ncols, size_t col_stride, u64 *digests);
/* nodes: [2*nleaves][5]; nodes[1] = root, nodes[nleaves + j] = leaf j */
void orc_merkle_build(const u64 *leaves, size_t nleaves, u64 *nodes);

int orc_num_threads(void);
void orc_set_num_threads(int n);   /* OpenMP team size of the parallel loops (no-op without OpenMP) */
#ifdef __cplusplus
}
#endif
/* xpoly.c: array stages of a complete prove (X-field arrays interleaved [k][3], Montgomery form) */
void orc_xpoly_eval(const u64 *c, size_t len, const u64 x[3], u64 out[3]);
void orc_bpoly_eval_x(const u64 *c, size_t len, const u64 x[3], u64 out[3]);
void orc_xpoly_fold(const u64 *c, size_t len, size_t ff, const u64 r[3], u64 *out);
void orc_xpoly_mul(const u64 *a, size_t la, const u64 *b, size_t lb, u64 *out);
void orc_xpoly_div(u64 *num, size_t ln, const u64 *den, size_t ld, u64 *q);
void orc_xzerofier(const u64 *pts, size_t k, u64 *out);
void orc_xinterpolate(const u64 *xs, const u64 *ys, size_t k, u64 *out);
void orc_xpoly_axpy(u64 *dst, const u64 *src, size_t len, const u64 w[3]);
void orc_xpoly_add_scaled_arg(u64 *dst, const u64 *src, size_t len, u64 scale, u64 arg);
void orc_interpolants_table(const u64 *trace, unsigned log2n, size_t ncols, const u64 *rand, size_t h, u64 *out);
void orc_bary_weights(unsigned log2n, const u64 alpha[3], u64 *dods, u64 denom_inv[3]);
void orc_ood_row(const u64 *cols, size_t ncols, unsigned log2n, int xf, const u64 *dods, const u64 denom_inv[3], const u64 *rand,
                 size_t h, const u64 alpha[3], u64 *out);
void orc_weighted_colsum(const u64 *main_coef, size_t nmain, const u64 *aux_coef, size_t naux, size_t len, const u64 *w, u64 *out);
void orc_deep_combination(const u64 *cw_ma, const u64 *cw_p, const u64 *cw_r, unsigned log2N, u64 offset, const u64 *points,
                          const u64 *values, const u64 *weights, u64 *out);
/* auxiliary-table extension from the AIR-derived rules (aux_extend.c); Montgomery form, column-major planes */
void orc_aux_extend(const u64 *main_t, size_t n, const u64 *ch, u64 *aux_t);
void orc_fill_derived_main(u64 *main_t, size_t n);
#endif

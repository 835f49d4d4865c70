/* CPU checker for the auxiliary-table extension (TEST INFRASTRUCTURE ONLY).
 *
 * Runs the rules generated from the AIR by triton-vm_b200/airgen/extend_gen.py (aux_extend_gen.inc) as the plain
 * sequential loop of MasterMainTable::extend (master_table.rs:1006-1075): per auxiliary column a running value that
 * each row updates through an affine map, then the 41 degree-lowering columns row by row (substitutions.rs:336-368).
 * The device (csrc/aux_extend.cu) evaluates the same rules as a parallel scan; tests compare both with the
 * AIR-solving restatement in oracle/tracegen.py.  All values Montgomery form; tables column-major, X-field columns as
 * three planes: aux[(3*col + plane) * n + row]. */
#include <stddef.h>
#include "field_inline.h"
#include "tvm_oracle.h"

typedef struct {
  u64 *main_t; u64 *aux_t; const u64 *ch; size_t n, cur, nxt;
} auxctx;

static inline xfe xlift(u64 b) { xfe r = {b, 0, 0}; return r; }
static inline u64 fneg(u64 a) { return fsub(0, a); }
static inline xfe xneg(xfe a) { xfe r = {fsub(0, a.c0), fsub(0, a.c1), fsub(0, a.c2)}; return r; }
static inline int xis_zero(xfe a) { return (a.c0 | a.c1 | a.c2) == 0; }
static inline xfe xinv(xfe a) { u64 in[3] = {a.c0, a.c1, a.c2}, out[3]; orc_xinv(in, out); xfe r = {out[0], out[1], out[2]}; return r; }
static inline xfe aux_load(const auxctx *c, int col, size_t row) {
  xfe r = {c->aux_t[(size_t)(3 * col) * c->n + row], c->aux_t[(size_t)(3 * col + 1) * c->n + row], c->aux_t[(size_t)(3 * col + 2) * c->n + row]};
  return r;
}
static inline void aux_store(const auxctx *c, int col, size_t row, xfe v) {
  c->aux_t[(size_t)(3 * col) * c->n + row] = v.c0;
  c->aux_t[(size_t)(3 * col + 1) * c->n + row] = v.c1;
  c->aux_t[(size_t)(3 * col + 2) * c->n + row] = v.c2;
}
static inline xfe ch_load(const auxctx *c, int i) { xfe r = {c->ch[3 * i], c->ch[3 * i + 1], c->ch[3 * i + 2]}; return r; }

#define AUXGEN_FN static
#define AUXGEN_ARGS const auxctx *c
#define AUXGEN_PASS c
#define AUXGEN_TOUCH (void)c
#define MC(col) (c->main_t[(size_t)(col) * c->n + c->cur])
#define MN(col) (c->main_t[(size_t)(col) * c->n + c->nxt])
#define AC(col) aux_load(c, (col), c->cur)
#define AN(col) aux_load(c, (col), c->nxt)
#define CH(i) ch_load(c, (i))
#define AW(col, v) aux_store(c, (col), c->cur, (v))
#define MW(col, v) (c->main_t[(size_t)(col) * c->n + c->cur] = (v))
#include "aux_extend_gen.inc"

/* main_t: [379][n]; ch: [63][3]; aux_t: [91*3][n], column 90 (the batch randomizer) is left as the caller filled it */
void orc_aux_extend(const u64 *main_t, size_t n, const u64 *ch, u64 *aux_t) {
  auxctx c = {(u64 *)main_t, aux_t, ch, n, 0, 0};
  for (int level = 0; level < AUXGEN_NUM_LEVELS; level++)
    for (int q = 0; q < AUXGEN_NUM_BASE; q++) {
      if (AUXGEN_LEVEL[q] != level) continue;
      xfe a, b, v = {0, 0, 0};
      c.cur = c.nxt = 0;
      if (auxgen_init(q, &c, &b)) v = b;
      aux_store(&c, q, 0, v);
      for (size_t i = 1; i < n; i++) {
        c.cur = i - 1; c.nxt = i;
        if (auxgen_tran(q, &c, &a, &b)) v = xadd(xmul(a, v), b);
        aux_store(&c, q, i, v);
      }
    }
  const xfe zero = {0, 0, 0};
  for (size_t i = 0; i + 1 < n; i++) {
    c.cur = i; c.nxt = i + 1;
    auxgen_derived_tran(&c);
  }
  c.cur = n - 1;
  for (int k = 0; k < AUXGEN_NUM_DERIVED_TRAN; k++) aux_store(&c, AUXGEN_DERIVED_START_TRAN + k, n - 1, zero);
}

/* DegreeLoweringTable::fill_derived_main_columns (substitutions.rs:128-161): columns 149.. of main_t [379][n] from 0..148 */
void orc_fill_derived_main(u64 *main_t, size_t n) {
  auxctx c = {main_t, 0, 0, n, 0, 0};
  for (size_t i = 0; i < n; i++) {
    c.cur = c.nxt = i;
    auxgen_derived_main_init(&c);
    auxgen_derived_main_cons(&c);
  }
  for (size_t i = 0; i + 1 < n; i++) {
    c.cur = i; c.nxt = i + 1;
    auxgen_derived_main_tran(&c);
  }
  for (int k = 0; k < AUXGEN_NUM_DERIVED_MAIN_TRAN; k++) main_t[(size_t)(AUXGEN_DERIVED_MAIN_START_TRAN + k) * n + n - 1] = 0;
}

/* Plain-C client of tvm_prove: the padded main table (the 149 table columns are enough), the randomness and the claim come
 * from files; degree-lowering columns, auxiliary table and proof are produced on the GPU; the proof is then checked with
 * tvm_verify.  This is the call sequence a patched `Prover::prove` performs, without Rust.
 *
 *   prove_tables <dir>     with <dir>/claim.txt   security log2_expansion ldt_choice padded_height  digest[5]  n_in in...  n_out out...
 *                               <dir>/main.u64    [379][n]      canonical little-endian u64, n = trace domain length
 *                               <dir>/main_rand.u64 [379][h]    trace-randomizer coefficients
 *                               <dir>/aux_rand.u64  [91][h][3]
 *                               <dir>/col90.u64     [n][3]      batch-randomizer column
 *                               <dir>/quot_rand.u64 [q][3]      quotient-segment randomizer
 *                          writes <dir>/proof.u64
 *
 *   gcc -std=c99 -Iinclude examples/prove_tables.c -Ltriton-vm_b200/lib -ltvm_b200 -o prove_tables
 *
 * Exit status: 0 proof written and verified, 1 verification failed, 2 usage / input error, 3 no usable GPU (there is no
 * CPU fallback), 4 the library reported an error. */
#define _POSIX_C_SOURCE 199309L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "tvm_b200.h"

static double now_ms(void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return 1e3 * (double)t.tv_sec + 1e-6 * (double)t.tv_nsec;
}

typedef struct {
  tvm_ctx *ctx;
  const uint64_t *main_trace, *col90;
  uint64_t *aux_trace, *aux_rand;
  unsigned log2_n;
} extend_state;

/* MasterMainTable::extend on the device, from inside the callback (same context: re-entrant) */
static int extend_on_device(void *user, const uint64_t *challenges, uint64_t **aux_trace, uint64_t **aux_rand) {
  extend_state *st = (extend_state *)user;
  int rc = tvm_aux_extend(st->ctx, st->main_trace, st->log2_n, challenges, st->col90, st->aux_trace);
  *aux_trace = st->aux_trace;
  *aux_rand = st->aux_rand;
  return rc;
}

static uint64_t *read_u64(const char *dir, const char *name, size_t count) {
  char path[4096];
  snprintf(path, sizeof path, "%s/%s", dir, name);
  FILE *f = fopen(path, "rb");
  uint64_t *buf = malloc((count ? count : 1) * sizeof *buf);
  if (!f || !buf || fread(buf, sizeof *buf, count, f) != count) { fprintf(stderr, "cannot read %zu words from %s\n", count, path); exit(2); }
  fclose(f);
  return buf;
}

int main(int argc, char **argv) {
  if (argc != 2) { fprintf(stderr, "usage: %s <dir>\n", argv[0]); return 2; }
  const char *dir = argv[1];
  char path[4096];
  snprintf(path, sizeof path, "%s/claim.txt", dir);
  FILE *f = fopen(path, "r");
  if (!f) { perror(path); return 2; }
  unsigned long long v[4], w;
  tvm_claim claim;
  memset(&claim, 0, sizeof claim);
  if (fscanf(f, "%llu %llu %llu %llu", &v[0], &v[1], &v[2], &v[3]) != 4) return 2;
  for (int i = 0; i < 5; i++) { if (fscanf(f, "%llu", &w) != 1) return 2; claim.program_digest[i] = w; }
  claim.version = 6; /* proof.rs:33 CURRENT_VERSION */
  uint64_t *io[2]; size_t nio[2];
  for (int k = 0; k < 2; k++) {
    if (fscanf(f, "%llu", &w) != 1) return 2;
    nio[k] = (size_t)w; io[k] = malloc((nio[k] + 1) * sizeof(uint64_t));
    for (size_t i = 0; i < nio[k]; i++) { if (fscanf(f, "%llu", &w) != 1) return 2; io[k][i] = w; }
  }
  fclose(f);
  claim.input = io[0]; claim.num_input = nio[0]; claim.output = io[1]; claim.num_output = nio[1];
  tvm_params params = {(uint32_t)v[0], (uint32_t)v[1], (uint32_t)v[2], 0};
  const uint64_t padded_height = v[3];

  tvm_domains dom;
  int rc = tvm_derive_domains(&params, padded_height, &dom);
  if (rc) { fprintf(stderr, "tvm_derive_domains: %s\n", tvm_strerror(rc)); return 4; }
  const size_t n = dom.trace_len, h = dom.num_trace_randomizers, q = dom.num_quotient_randomizer_coefficients;
  unsigned log2_n = 0;
  while (((size_t)1 << log2_n) < n) log2_n++;

  tvm_ctx *ctx = NULL;
  rc = tvm_ctx_create(&ctx, 0);
  if (rc) { fprintf(stderr, "tvm_ctx_create: %s\n", tvm_strerror(rc)); return 3; }

  uint64_t *main_trace = read_u64(dir, "main.u64", (size_t)TVM_NUM_MAIN_COLUMNS * n);
  uint64_t *main_rand = read_u64(dir, "main_rand.u64", (size_t)TVM_NUM_MAIN_COLUMNS * h);
  uint64_t *aux_rand = read_u64(dir, "aux_rand.u64", (size_t)TVM_NUM_AUX_COLUMNS * h * 3);
  uint64_t *col90 = read_u64(dir, "col90.u64", n * 3);
  uint64_t *quot_rand = read_u64(dir, "quot_rand.u64", q * 3);
  uint64_t *aux_trace = malloc((size_t)TVM_NUM_AUX_COLUMNS * n * 3 * sizeof *aux_trace);
  size_t proof_len = (size_t)1 << 22;
  uint64_t *proof = malloc(proof_len * sizeof *proof);
  if (!aux_trace || !proof) return 2;

  /* PROVE_TABLES_REPS > 1 repeats the whole sequence (same inputs, same proof) so that the last, warm repetition can be read
   * as a wall-clock figure for "149 table columns in host memory -> proof in host memory" */
  const char *reps_env = getenv("PROVE_TABLES_REPS");
  const int reps = reps_env && atoi(reps_env) > 0 ? atoi(reps_env) : 1;
  const size_t proof_cap = proof_len;
  for (int rep = 0; rep < reps; rep++) {
    const double t0 = now_ms();
    rc = tvm_fill_derived_main_columns(ctx, main_trace, log2_n);            /* columns 149..378 */
    if (rc) { fprintf(stderr, "tvm_fill_derived_main_columns: %s (%s)\n", tvm_strerror(rc), tvm_last_error(ctx)); return 4; }
    const double t1 = now_ms();
    extend_state st = {ctx, main_trace, col90, aux_trace, aux_rand, log2_n};
    proof_len = proof_cap;
    rc = tvm_prove(ctx, &params, &claim, padded_height, main_trace, main_rand, extend_on_device, &st, quot_rand, proof, &proof_len);
    if (rc) { fprintf(stderr, "tvm_prove: %s (%s)\n", tvm_strerror(rc), tvm_last_error(ctx)); return 4; }
    printf("repetition %d: derived main columns %.1f ms, prove incl. extend on the device %.1f ms\n", rep, t1 - t0, now_ms() - t1);
  }
  tvm_ctx_destroy(ctx);

  char why[256];
  rc = tvm_verify(&params, &claim, proof, proof_len, 0, why, sizeof why);
  snprintf(path, sizeof path, "%s/proof.u64", dir);
  f = fopen(path, "wb");
  if (!f || fwrite(proof, sizeof *proof, proof_len, f) != proof_len) { perror(path); return 2; }
  fclose(f);
  printf("proof: %zu words, %s\n", proof_len, rc == TVM_OK ? "verified" : why);
  return rc == TVM_OK ? 0 : 1;
}

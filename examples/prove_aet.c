/* Plain-C client of tvm_prove_aet: the AlgebraicExecutionTrace (the arrays the reference's `aet.rs:41-91` holds), the
 * randomness and the claim come from files; MasterMainTable::new + pad, the degree-lowering columns, the auxiliary table and
 * the proof are produced on the GPU; the proof is then checked with tvm_verify (AIR included).  What a patched
 * `Prover::prove` does when it hands the witness itself to the library, without Rust.
 *
 *   prove_aet <dir>     <dir>/claim.txt      security log2_expansion ldt_choice padded_height  digest[5]  n_in in...  n_out out...
 *                       <dir>/aet_<field>.u64 (aet_instruction_multiplicities.u32): one file per field of tvm_aet, row-major,
 *                                            canonical little-endian words; row counts follow from the file sizes
 *                       <dir>/main_rand.u64 [379][h], aux_rand.u64 [91][h][3], col90.u64 [n][3], quot_rand.u64 [q][3]
 *                       (tools/make_workload.py <workload> <dir> --aet writes all of them)
 *                       writes <dir>/proof_aet.u64
 *
 *   gcc -std=c99 -Iinclude examples/prove_aet.c -Ltriton-vm_b200/lib -ltvm_b200 -o prove_aet
 *
 * Exit status: 0 proof written and verified, 1 verification failed, 2 usage / input error, 3 no usable GPU (there is no
 * CPU fallback), 4 the library reported an error. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "tvm_b200.h"

/* whole file -> malloc'ed buffer; *count = number of `elem`-byte elements */
static void *read_all(const char *dir, const char *name, size_t elem, uint64_t *count) {
  char path[4096];
  snprintf(path, sizeof path, "%s/%s", dir, name);
  FILE *f = fopen(path, "rb");
  if (!f) { perror(path); exit(2); }
  fseek(f, 0, SEEK_END);
  const long bytes = ftell(f);
  fseek(f, 0, SEEK_SET);
  void *buf = malloc(bytes > 0 ? (size_t)bytes : 1);
  if (bytes < 0 || !buf || (size_t)bytes % elem || fread(buf, 1, (size_t)bytes, f) != (size_t)bytes) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
  fclose(f);
  *count = (uint64_t)((size_t)bytes / elem);
  return buf;
}

static uint64_t *read_rows(const char *dir, const char *name, unsigned width, uint64_t *rows) {
  uint64_t words;
  uint64_t *p = read_all(dir, name, sizeof(uint64_t), &words);
  if (words % width) { fprintf(stderr, "%s: not a multiple of %u words\n", name, width); exit(2); }
  *rows = words / width;
  return p;
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

  tvm_aet aet;
  memset(&aet, 0, sizeof aet);
  uint64_t count;
  aet.program = read_all(dir, "aet_program.u64", 8, &aet.program_len);
  aet.instruction_multiplicities = read_all(dir, "aet_instruction_multiplicities.u32", 4, &count);
  if (count != aet.program_len) { fprintf(stderr, "one multiplicity per program word expected\n"); return 2; }
  aet.processor_trace = read_rows(dir, "aet_processor_trace.u64", 39, &aet.processor_rows);
  aet.op_stack_underflow_trace = read_rows(dir, "aet_op_stack_underflow_trace.u64", 4, &aet.op_stack_rows);
  aet.ram_trace = read_rows(dir, "aet_ram_trace.u64", 7, &aet.ram_rows);
  aet.program_hash_trace = read_rows(dir, "aet_program_hash_trace.u64", 67, &aet.program_hash_rows);
  aet.sponge_trace = read_rows(dir, "aet_sponge_trace.u64", 67, &aet.sponge_rows);
  aet.hash_trace = read_rows(dir, "aet_hash_trace.u64", 67, &aet.hash_rows);
  aet.u32_entries = read_rows(dir, "aet_u32_entries.u64", 4, &aet.u32_count);
  aet.cascade_table_lookup_multiplicities = read_rows(dir, "aet_cascade_table_lookup_multiplicities.u64", 2, &aet.cascade_count);
  aet.lookup_table_lookup_multiplicities = read_rows(dir, "aet_lookup_table_lookup_multiplicities.u64", 256, &count);
  if (count != 1) { fprintf(stderr, "256 lookup multiplicities expected\n"); return 2; }

  tvm_domains dom;
  int rc = tvm_derive_domains(&params, padded_height, &dom);
  if (rc) { fprintf(stderr, "tvm_derive_domains: %s\n", tvm_strerror(rc)); return 4; }
  const uint64_t n = dom.trace_len, h = dom.num_trace_randomizers, q = dom.num_quotient_randomizer_coefficients;
  uint64_t *main_rand = read_all(dir, "main_rand.u64", 8, &count);
  if (count != (uint64_t)TVM_NUM_MAIN_COLUMNS * h) { fprintf(stderr, "main_rand.u64: [379][%llu] expected\n", (unsigned long long)h); return 2; }
  uint64_t *aux_rand = read_all(dir, "aux_rand.u64", 8, &count);
  if (count != (uint64_t)TVM_NUM_AUX_COLUMNS * h * 3) { fprintf(stderr, "aux_rand.u64: [91][%llu][3] expected\n", (unsigned long long)h); return 2; }
  uint64_t *col90 = read_all(dir, "col90.u64", 8, &count);
  if (count != n * 3) { fprintf(stderr, "col90.u64: [%llu][3] expected\n", (unsigned long long)n); return 2; }
  uint64_t *quot_rand = read_all(dir, "quot_rand.u64", 8, &count);
  if (count != q * 3) { fprintf(stderr, "quot_rand.u64: [%llu][3] expected\n", (unsigned long long)q); return 2; }

  tvm_ctx *ctx = NULL;
  rc = tvm_ctx_create(&ctx, 0);
  if (rc) { fprintf(stderr, "tvm_ctx_create: %s\n", tvm_strerror(rc)); return 3; }
  size_t proof_len = (size_t)1 << 22;
  uint64_t *proof = malloc(proof_len * sizeof *proof);
  if (!proof) return 2;
  rc = tvm_prove_aet(ctx, &params, &claim, padded_height, &aet, main_rand, aux_rand, col90, quot_rand, proof, &proof_len);
  if (rc) { fprintf(stderr, "tvm_prove_aet: %s (%s)\n", tvm_strerror(rc), tvm_last_error(ctx)); return 4; }
  tvm_ctx_destroy(ctx);

  char why[256];
  rc = tvm_verify(&params, &claim, proof, proof_len, 0, why, sizeof why);
  snprintf(path, sizeof path, "%s/proof_aet.u64", dir);
  f = fopen(path, "wb");
  if (!f || fwrite(proof, sizeof *proof, proof_len, f) != proof_len) { perror(path); return 2; }
  fclose(f);
  printf("proof from the AET (%llu processor rows, %llu RAM rows): %zu words, %s\n", (unsigned long long)aet.processor_rows,
         (unsigned long long)aet.ram_rows, proof_len, rc == TVM_OK ? "verified" : why);
  return rc == TVM_OK ? 0 : 1;
}

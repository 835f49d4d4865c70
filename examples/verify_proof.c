/* Plain-C client of libtvm_b200: reads a claim and a proof from a text file of decimal words and runs Stark::verify.
 *
 *   file layout: security_level log2_expansion ldt_choice version
 *                program_digest[5]  num_input input...  num_output output...  proof_len proof...
 *
 *   gcc -std=c99 -Iinclude examples/verify_proof.c -Ltriton-vm_b200/lib -ltvm_b200 -o verify_proof
 *
 * Shows that include/tvm_b200.h is a C header (no C++ types cross the boundary) and that the verifier needs neither a
 * context nor a GPU.  Exit status 0 = accepted, 1 = rejected (reason on stderr), 2 = usage / input error. */
#include <stdio.h>
#include <stdlib.h>
#include "tvm_b200.h"

static int read_words(FILE *f, uint64_t *out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    unsigned long long v;
    if (fscanf(f, "%llu", &v) != 1) return 0;
    out[i] = (uint64_t)v;
  }
  return 1;
}

int main(int argc, char **argv) {
  if (argc != 2) { fprintf(stderr, "usage: %s <claim-and-proof.txt>\n", argv[0]); return 2; }
  FILE *f = fopen(argv[1], "r");
  if (!f) { perror(argv[1]); return 2; }
  uint64_t head[4], count;
  tvm_claim claim;
  if (!read_words(f, head, 4) || !read_words(f, claim.program_digest, 5) || !read_words(f, &count, 1)) return 2;
  tvm_params params = {(uint32_t)head[0], (uint32_t)head[1], (uint32_t)head[2], 0};
  claim.version = (uint32_t)head[3];
  uint64_t *input = malloc((count + 1) * sizeof *input);
  if (!input || !read_words(f, input, count)) return 2;
  claim.input = input; claim.num_input = count;
  if (!read_words(f, &count, 1)) return 2;
  uint64_t *output = malloc((count + 1) * sizeof *output);
  if (!output || !read_words(f, output, count)) return 2;
  claim.output = output; claim.num_output = count;
  if (!read_words(f, &count, 1)) return 2;
  uint64_t *proof = malloc((count + 1) * sizeof *proof);
  if (!proof || !read_words(f, proof, count)) return 2;
  fclose(f);

  uint64_t padded_height = 0;
  char why[256];
  if (tvm_proof_padded_height(proof, count, &padded_height) != TVM_OK) { fprintf(stderr, "not a proof\n"); return 1; }
  int rc = tvm_verify(&params, &claim, proof, count, 0, why, sizeof why);
  if (rc == TVM_OK) printf("accepted (padded height %llu, %llu words)\n", (unsigned long long)padded_height, (unsigned long long)count);
  else fprintf(stderr, "rejected: %s (%s)\n", why, tvm_strerror(rc));
  free(input); free(output); free(proof);
  return rc == TVM_OK ? 0 : 1;
}

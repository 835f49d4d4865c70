// Prover-level declarations shared by prover.cu, stark_kernels.cu and the C ABI.
#pragma once
#include <string>
#include <vector>
#include "ctx.h"

namespace tvm {

static constexpr int NUM_QUOTIENT_SEGMENTS = 4;              // stark.rs:66 (= air::TARGET_DEGREE)
static constexpr int NUM_RANDOMIZED_QUOTIENT_SEGMENTS = 5;   // stark.rs:75

struct StarkParams {          // stark.rs:113-145 (proven regime)
  unsigned security_level;
  unsigned log2_expansion;
  unsigned ldt_choice;        // 0 = heuristic (stark.rs:1942-1957), 1 = FRI, 2 = STIR
  unsigned soundness;         // 0 = ProximityRegime::Proven (default), 1 = ::Conjectured
};
static constexpr int STIR_MAX_ROUNDS = 16;
static constexpr unsigned STIR_LOG2_FOLDING_FACTOR = 2;      // stark.rs:2023
struct StirDerived {          // stir.rs:112-147
  size_t folding_factor;
  int num_rounds;
  size_t in_domain[STIR_MAX_ROUNDS], out_of_domain[STIR_MAX_ROUNDS];
  size_t final_num_in_domain_queries, final_degree;
};
struct StarkDerived {
  size_t padded_height, num_trace_randomizers, randomized_trace_len, trace_len, quotient_len, ldt_len;
  u64 ldt_offset;  // canonical
  size_t num_collinearity_checks, fri_num_rounds, fri_last_round_max_degree, num_quotient_randomizer_coefficients;
  int ldt;                    // 1 = FRI, 2 = STIR
  size_t num_first_round_queries;
  StirDerived stir;
};
int stir_derive(unsigned security_level, unsigned log2_folding_factor, unsigned log2_initial_expansion, unsigned log2_high_degree_bound,
                bool conjectured, StirDerived &out);
int stark_derive(const StarkParams &sp, size_t padded_height, StarkDerived &d);

struct ClaimView {            // proof.rs:68-88, canonical words
  const u64 *program_digest;  // 5
  unsigned version;
  const u64 *input; size_t num_input;
  const u64 *output; size_t num_output;
};
// Given the 63 challenges (canonical), provides aux_trace [91][n][3] and aux_rand [91][h][3] (canonical).
// The callee sets *aux_trace / *aux_rand to its own (ideally pinned) buffers, which must stay valid
// until prove returns (zero copy: the library reads them straight into the H2D transfer).
typedef int (*AuxCallback)(void *user, const u64 *challenges, u64 **aux_trace, u64 **aux_rand);

struct ProveTimings {
  std::vector<std::pair<std::string, float>> stages;   // device ms per stage, reference profiler labels
  bool low_memory = false;                             // the last prove() ran in low-memory (just-in-time LDE) mode
};

// Table stages on the device (tvm_prove_tables): MasterMainTable::extend (master_table.rs:1006-1075) runs inside the prove
// on the resident main trace instead of in a host callback; optionally the 230 degree-lowering main columns
// (fill_derived_main_columns, substitutions.rs:128-161) are computed on the device too and never uploaded.
struct DeviceTables {
  const u64 *aux_rand;            // [91][h][3] canonical (host or device)
  const u64 *randomizer_column;   // [n][3] canonical: aux column 90 (master_table.rs:1019-1025); nullptr = zeros
  bool fill_derived_main;         // columns 149..378 of the main trace are computed, not read
  const tvm_aet *aet = nullptr;   // the nine tables' columns come from the AET (main_fill.cu), no main trace is read at all
};

void stark_prove(Ctx &c, const StarkParams &sp, const ClaimView &claim, size_t padded_height, const u64 *h_main_trace,
                 const u64 *h_main_rand, AuxCallback aux_cb, void *aux_user, const u64 *h_quot_rand, std::vector<u64> &proof,
                 ProveTimings *timings, const DeviceTables *dev_tables = nullptr, const struct ExternalTranscript *ext = nullptr);

// ---- main_fill.cu: MasterMainTable::new + pad from the AET; d_table [>= 149][n] receives CANONICAL columns ----
void main_fill_run(Ctx &c, struct DevMem &mem, const tvm_aet &aet, size_t n, u64 *d_table, uint64_t *lengths9);
void bezout_run(Ctx &c, struct DevMem &mem, const u64 *d_roots_mont, size_t m, u64 *d_a, u64 *d_b);

// ---- stark_kernels.cu ----
struct SegmentArgs {
  const u64 *quot; size_t quot_stride;
  const u64 *rnd; size_t rnd_stride; unsigned rnd_len;
  u64 *out; size_t out_stride;
  size_t seg_len;
  PowTab zeta4, off;
  u64 zeta_pow[4];
};
struct DeepArgs {
  const u64 *cw; size_t cw_stride;
  u64 *out; size_t out_stride;
  int log_n, log_r;                          // trace length 2^log_n, domain of 2^(log_n+log_r) points
  unsigned coset_first, coset_step, num_cosets;   // local coset y is domain coset coset_first + coset_step*y
  PowTab dom;
  u64 offset;
  xfe point[4], value[4], weight[4];
};
void shards_to_natural_run(Ctx &c, const u64 *in, u64 *out, size_t rank_stride, size_t plane_stride, size_t out_stride, int log_n, int log_r,
                           int log_w, int planes);
void shard_digests_to_natural_run(Ctx &c, const u64 *in, u64 *out, int log_n, int log_r, int log_w);
void coset_to_natural_run(Ctx &c, const u64 *in, u64 *out, size_t in_stride, size_t out_stride, int log_n, int log_r, int planes);
void deinterleave3_run(Ctx &c, const u64 *in, u64 *out, size_t len, size_t ncols);
void segment_chain_run(Ctx &c, const SegmentArgs &a);
void xpow_vector_run(Ctx &c, xfe base, u64 *out, size_t stride, size_t len);
void col_dot_run(Ctx &c, const u64 *cols, size_t col_stride, size_t ncols, size_t len, const u64 *xvec, size_t xvec_stride,
                 size_t xvec_set_stride, int nvec, u64 *out);
void weighted_colsum_run(Ctx &c, const u64 *cols, size_t col_stride, unsigned ncols, bool xfield, const u64 *d_w, size_t len, u64 *out,
                         size_t out_stride, bool accumulate);
void deep_run(Ctx &c, const DeepArgs &a);
void fri_leaves_run(Ctx &c, const u64 *cw, size_t stride, size_t n, u64 *leaves);
void fri_fold_run(Ctx &c, const u64 *in, size_t in_stride, size_t n, u64 offset_mont, xfe chal, u64 *out, size_t out_stride);
void gather_rows_run(Ctx &c, const u64 *table, size_t col_stride, unsigned ncols, const unsigned *d_idx, unsigned nidx, int log_n, int log_r,
                     u64 *d_out, int log_w = 0, unsigned rank = 0, unsigned coset_mem_stride = 1);
void gather_rows_scatter_run(Ctx &c, const u64 *table, size_t col_stride, unsigned ncols, const unsigned *d_kt, unsigned count, u64 *d_out);
void gather_digests_run(Ctx &c, const u64 *nodes, const unsigned *d_idx, unsigned nidx, u64 *d_out, int log_w = 0, unsigned rank = 0);
void scale_by_powers_run(Ctx &c, u64 *v, size_t stride, int planes, size_t len, PowTab tab);

}  // namespace tvm

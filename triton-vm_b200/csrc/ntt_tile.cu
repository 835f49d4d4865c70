// Round-2 NTT kernels: square two-round tiles with register-resident radix-R DFTs (R = 8, 16, 32).
//
// A transform of size n = n2 * n1 (reference semantics: ntt.cu header; call sites stark.rs:872-877,
// arithmetic_domain.rs:141-212, master_table.rs:258-322) is two passes over HBM/L2:
//   pass A: n1/T tiles; a tile is T adjacent strided rows (fixed j1, elements j = j1 + n1 j2) -> n2-point NTT over j2,
//           times the inter-pass twist, stored transposed (tmp[K n1 + j1]);
//   pass B: n2/T tiles of T contiguous rows of tmp -> n1-point NTT over j1, stored to natural order (k = K + n2 K_B).
// A tile row of M = R*R points is transformed in TWO rounds: thread (t, a) holds the R points j = a + R b in registers,
// does an R-point DFT whose twiddles are all powers of two (ntt_radix.cuh), multiplies by w_M^(a k2) and hands the
// results over through shared memory; thread (t, k2) then does the R-point DFT over a and owns X[k2 + R k1].  Per point
// and pass: one generic twiddle multiplication, one shared-memory round trip, no bit reversal (the register index
// permutation is free), against 3.3 rounds, 2.6 twiddles and the bit-reversed addressing of the radix-8 tiles in ntt.cu.
// The row index t occupies the low lane bits, so every warp-wide global access covers whole 32/64-byte runs of adjacent
// rows and the exchange buffer ((k2 (R+1) + a) T + t) is bank-conflict free in both directions.
//
// Pre-/post-operations fused into the passes (each costs what its arithmetic costs, no extra pass over memory):
//   * coset evaluation (LDE): pass A multiplies point j2 by S_c[j2] = w_{r n2}^(c j2) (a table of n2 entries per
//     coset) and folds the randomizer chunk (c'_j += w_r^c c'_{n+j}, j < fold_count); the factor w_{rn}^(c j1) that
//     remains of the coset shift is one more term in the twist exponent: twist = w_{rn}^(j1 (r K + c)).
//   * interpolation: pass B multiplies by n^-1 offset^k (the coefficients are kept pre-scaled by offset^k).
//   Both are geometric in k1 for a fixed thread: factor = base * step^k1, two table look-ups per thread and tile.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include "ctx.h"
#include "launch.h"
#include "ntt_radix.cuh"

namespace tvm {

struct TileJob {
  const u64 *in;
  u64 *out;
  size_t in_col_stride, in_coset_stride, out_col_stride, out_coset_stride;   // per blockIdx.z (column), blockIdx.y (coset)
  size_t in_row_stride, in_elem_stride, out_row_stride, out_elem_stride;     // element idx of row r at r*row + idx*elem
  const u64 *tw;             // w_M^(+-e), e < M = R*R
  // pre-operation (coset evaluation)
  const u64 *prescale;       // [gridDim.y][M] or nullptr
  const u64 *fold_factor;    // [gridDim.y]
  unsigned fold_count;       // input elements at offset < fold_count receive + fold_factor * in[fold_offset + offset]
  size_t fold_offset;
  // post-operation: 0 none; 1 twist G^(row (mul K + c_y)), c_y = coset_first + coset_step y; 2 scalar * G^(row + mul K)
  int post_mode;
  PowTab G;
  u64 exp_mask;
  u64 mul;
  unsigned coset_first, coset_step;
  u64 scalar;
};

__device__ __forceinline__ u64 powtab_at(const PowTab &t, u64 e) {
  const u64 lo = __ldg(t.lo + (e & ((1ULL << t.shift) - 1)));
  const u64 hi = __ldg(t.hi + (e >> t.shift));
  return fmul(lo, hi);
}

static constexpr int TILE_THREADS = 256;

// TMA bulk copies (cp.async.bulk, SASS UBLKCP) with mbarrier completion - used to stage the contiguous rows of a pass-B tile
__device__ __forceinline__ unsigned tile_smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tile_mbar_init(u64 *bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n\tfence.mbarrier_init.release.cluster;" ::"r"(tile_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tile_mbar_expect_tx(u64 *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tile_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tile_bulk_g2s(void *dst, const void *src, unsigned bytes, u64 *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tile_smem_addr(dst)), "l"(src),
               "r"(bytes), "r"(tile_smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void tile_mbar_wait(u64 *bar, unsigned parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "TVM_TILE_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra TVM_TILE_DONE;\n\t"
      "bra TVM_TILE_WAIT;\n\t"
      "TVM_TILE_DONE:\n\t"
      "}" ::"r"(tile_smem_addr(bar)),
      "r"(parity)
      : "memory");
}

// PRE: coset pre-scale + randomizer fold on the way in; POST: 0 none, 1 twist, 2 scalar * power on the way out (compile-time:
// the unused variants would otherwise double the unrolled code, ~100 KB of SASS, past the instruction cache)
// STREAM: second pass - its input is the intermediate that should live in L2 only (last use: ld.global.cs = evict first) and
// its output is not read again by the LDE (st.global.cs), so that the NEXT column's intermediate finds room in the L2.
// BULK: the tile's rows are contiguous in memory (second pass): thread 0 hands the T rows to the TMA engine (one bulk copy of
// M words per row into the exchange buffer's memory, row pitch M + 2 words: bank-conflict free for the (t, a) read-out), the
// CTA waits on the mbarrier and fills its registers from shared memory; no lane issues a global load.
template <int LOGR, bool INV, bool PRE, int POST, bool LOOPED, bool STREAM, bool BULK>
__global__ void __launch_bounds__(TILE_THREADS, 2) ntt_tile_kernel(TileJob p) {
  constexpr int R = 1 << LOGR, M = R * R, LOGT = 8 - LOGR, T = 1 << LOGT;
  extern __shared__ __align__(16) u64 smem[];
  u64 *tw = smem;            // [M]
  u64 *ex = smem + M;        // [(k2 (R+1) + a) T + t]
  const int tid = threadIdx.x, t = tid & (T - 1), a = tid >> LOGT;
  const size_t row = (size_t)blockIdx.x * T + t;
  const unsigned y = blockIdx.y;
  const u64 *in = p.in + (size_t)blockIdx.z * p.in_col_stride + (size_t)y * p.in_coset_stride;
  u64 *out = p.out + (size_t)blockIdx.z * p.out_col_stride + (size_t)y * p.out_coset_stride;
  __shared__ __align__(8) u64 bulk_bar;
  if (BULK) {
    if (tid == 0) tile_mbar_init(&bulk_bar);
    __syncthreads();
    if (tid == 0) {
      tile_mbar_expect_tx(&bulk_bar, (unsigned)(T * M * sizeof(u64)));
      const u64 *src = in + ((size_t)blockIdx.x * T) * p.in_row_stride;
      for (int r = 0; r < T; r++) tile_bulk_g2s(ex + (size_t)r * (M + 2), src + (size_t)r * p.in_row_stride, (unsigned)(M * sizeof(u64)), &bulk_bar);
    }
  }
  for (int i = tid; i < M; i += TILE_THREADS) tw[i] = __ldg(p.tw + i);

  // LOOPED: both rounds run through ONE copy of the unrolled R-point DFT (`unroll 1`), which keeps the kernel within the
  // instruction cache (ncu on the fully unrolled version: 97 KB of SASS for pass A, `stalled_no_instruction` 0.75 per
  // issued instruction) at the price of ~50 spilled words per thread; !LOOPED: two specialised copies, no spills.
  u64 v[R];
#pragma unroll(LOOPED ? 1 : 2)
  for (int rnd = 0; rnd < 2; rnd++) {
    if (rnd == 0) {
      const size_t off0 = row * p.in_row_stride + (size_t)a * p.in_elem_stride;
      const size_t step = (size_t)R * p.in_elem_stride;
      if (BULK) {
        tile_mbar_wait(&bulk_bar, 0);
        const u64 *st = ex + (size_t)t * (M + 2) + a;
#pragma unroll
        for (int b = 0; b < R; b++) v[b] = st[b * R];
      } else {
#pragma unroll
        for (int b = 0; b < R; b++) v[b] = STREAM ? __ldcs(in + off0 + b * step) : in[off0 + b * step];
      }
      if (PRE) {
        if (off0 < p.fold_count) v[0] = fadd(v[0], fmul(__ldg(p.fold_factor + y), in[p.fold_offset + off0]));
        const u64 *S = p.prescale + (size_t)y * M + a;
#pragma unroll
        for (int b = 0; b < R; b++) v[b] = fmul(v[b], __ldg(S + b * R));
      }
    } else {
      // second round: this thread owns outputs K = k2 + R k1 of row t, k2 = a
#pragma unroll
      for (int i = 0; i < R; i++) v[i] = ex[(a * (R + 1) + i) * T + t];
    }
    dft_pow2<LOGR, INV>(v);
    if (rnd == 0) {
      __syncthreads();           // tw[] complete
#pragma unroll
      for (int k2 = 0; k2 < R; k2++) {
        u64 z = v[bitrev_c(k2, LOGR)];
        if (k2) z = fmul(z, tw[a * k2]);          // a == 0: tw[0] = 1 (kept branch-free)
        ex[(k2 * (R + 1) + a) * T + t] = z;
      }
      __syncthreads();
    } else {
      const int k2 = a;
      const size_t ooff0 = row * p.out_row_stride + (size_t)k2 * p.out_elem_stride;
      const size_t ostep = (size_t)R * p.out_elem_stride;
      if (POST == 0) {
#pragma unroll
        for (int k1 = 0; k1 < R; k1++) {
          if (STREAM) __stcs(out + ooff0 + k1 * ostep, v[bitrev_c(k1, LOGR)]);
          else out[ooff0 + k1 * ostep] = v[bitrev_c(k1, LOGR)];
        }
      } else {
        u64 f, st;
        if (POST == 1) {
          const u64 c = p.coset_first + p.coset_step * y;
          f = powtab_at(p.G, (row * (p.mul * k2 + c)) & p.exp_mask);
          st = powtab_at(p.G, (row * p.mul * R) & p.exp_mask);
        } else {
          f = fmul(p.scalar, powtab_at(p.G, row + p.mul * k2));
          st = powtab_at(p.G, p.mul * R);
        }
#pragma unroll
        for (int k1 = 0; k1 < R; k1++) {
          const u64 o = fmul(v[bitrev_c(k1, LOGR)], f);
          if (STREAM) __stcs(out + ooff0 + k1 * ostep, o);
          else out[ooff0 + k1 * ostep] = o;
          if (k1 + 1 < R) f = fmul(f, st);
        }
      }
    }
  }
}

// randomizer fold-in after the interpolation (master_table.rs:392-403: interpolant + zerofier * randomizer, zerofier
// X^n - 1 on the trace domain): coef[k] -= rand[k] post^k, coef[n + k] = rand[k] post^(n + k), zero up to rand_pad.
__global__ void lde_randomizer_kernel(u64 *coef, size_t coef_stride, size_t n, const u64 *rand, unsigned rand_count,
                                      unsigned rand_pad, PowTab post) {
  const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= rand_pad) return;
  u64 *c = coef + (size_t)blockIdx.y * coef_stride;
  if (k < rand_count) {
    const u64 r = rand[(size_t)blockIdx.y * rand_count + k];
    c[k] = fsub(c[k], fmul(r, powtab_at(post, k)));
    c[n + k] = fmul(r, powtab_at(post, n + k));
  } else {
    c[n + k] = 0;
  }
}

__global__ void prescale_table_kernel(u64 *S, u64 *ff, PowTab pre, unsigned log_n, unsigned log_n1, unsigned M, unsigned coset_first,
                                      unsigned coset_step, u64 exp_mask) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  const u64 c = coset_first + coset_step * y;
  if (i < M) S[(size_t)y * M + i] = powtab_at(pre, ((c << log_n1) * i) & exp_mask);
  if (i == 0) ff[y] = powtab_at(pre, (c << log_n) & exp_mask);
}

u64 *Ctx::get_prescale(unsigned log_n, unsigned log_r, unsigned log_n1, unsigned M, unsigned first, unsigned step, unsigned count,
                        const std::function<void(u64 *, u64 *)> &fill) {
  auto key = std::make_tuple(log_n, log_r, log_n1, M, first, step, count);
  auto it = prescale_tabs.find(key);
  if (it != prescale_tabs.end()) return it->second;
  u64 *d = (u64 *)alloc(((size_t)count * M + count) * sizeof(u64));
  fill(d, d + (size_t)count * M);
  prescale_tabs[key] = d;
  return d;
}

static int tile_logr(int log_m) { return (log_m == 6 || log_m == 8 || log_m == 10) ? log_m / 2 : 0; }

template <bool INV, bool PRE, int POST, bool STREAM>
static void launch_tile(int logr, dim3 grid, cudaStream_t s, const TileJob &j) {
  const int R = 1 << logr, M = R * R, T = TILE_THREADS / R;
  const size_t smem = ((size_t)M + (size_t)R * (R + 1) * T) * sizeof(u64);
  auto go = [&](auto kernel) {
    TVM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<grid, TILE_THREADS, smem, s>>>(j);
  };
  static const bool looped = getenv("TVM_NTT_LOOPED_ROUNDS") != nullptr;   // A/B switch (measured 10 % slower at 2^20: spills)
  static const bool no_bulk = getenv("TVM_NTT_NO_TMA") != nullptr;         // A/B switch: per-lane loads in the second pass too
  // contiguous, 16-byte aligned rows of exactly M words (every second pass): staged by the TMA engine
  const bool bulk = !PRE && !no_bulk && j.in_elem_stride == 1 && j.in_row_stride == (size_t)M && ((uintptr_t)j.in % 16) == 0 &&
                    (j.in_col_stride * 8) % 16 == 0 && (j.in_coset_stride * 8) % 16 == 0;
  if (logr == 5 && looped) go(ntt_tile_kernel<5, INV, PRE, POST, true, STREAM, false>);
  else if (logr == 5 && bulk) go(ntt_tile_kernel<5, INV, false, POST, false, STREAM, true>);
  else if (logr == 5) go(ntt_tile_kernel<5, INV, PRE, POST, false, STREAM, false>);
  else if (logr == 4 && bulk) go(ntt_tile_kernel<4, INV, false, POST, false, STREAM, true>);
  else if (logr == 4) go(ntt_tile_kernel<4, INV, PRE, POST, false, STREAM, false>);
  else if (bulk) go(ntt_tile_kernel<3, INV, false, POST, false, STREAM, true>);
  else go(ntt_tile_kernel<3, INV, PRE, POST, false, STREAM, false>);
  TVM_CUDA(cudaGetLastError());
}

// bytes of transposed intermediate one launch pair may produce: pass B should find it in the L2 (126 MB, shared with the
// streaming output).  Measured DRAM traffic per column at 2^20, 8 cosets (ncu, caches left alone, no replay;
// profiles/r02h_ntt_tile_traffic.md): 157 MB with 64 MB pairs, 137 MB with 32 MB pairs on two streams (algorithmic: 72 MB;
// the L2 writes the dirty intermediate back once in any case: 8 + 64 + 64 = 136 MB is the floor of a two-pass transform).
static size_t tile_tmp_budget() {
  static const size_t mb = getenv("TVM_NTT_TMP_MB") ? (size_t)atoi(getenv("TVM_NTT_TMP_MB")) : 32;
  return std::max<size_t>(1, mb) << 20;
}
// launch pairs alternate between the context's stream and a second one (each with its own half of the intermediate buffer):
// a pair is 1024 CTAs = 3.46 waves of the 296 resident CTAs, the other stream's CTAs fill the partial last wave
static int tile_streams() {
  static const int k = getenv("TVM_NTT_STREAMS") ? atoi(getenv("TVM_NTT_STREAMS")) : 2;
  return k >= 2 ? 2 : 1;
}

// Picks n = n2 * n1 with both factors in {2^6, 2^8, 2^10}; false if this size is left to ntt.cu's kernels.
bool ntt_tile_supported(int log_n, int *log_n2, int *log_n1) {
  static const bool off = getenv("TVM_NTT_RADIX8_TILES") != nullptr;   // A/B: the round-1 kernels
  if (off) return false;
  for (int la = 10; la >= 6; la -= 2) {
    const int lb = log_n - la;
    if (lb > la) return false;
    if (lb == 6 || lb == 8 || lb == 10) { *log_n2 = la; *log_n1 = lb; return true; }
  }
  return false;
}

// Coset evaluation of pre-scaled coefficient columns (lde_evaluate_run's contract).
bool lde_evaluate_tiles(Ctx &c, const u64 *d_coef, size_t coef_stride, unsigned fold_count, unsigned log_n, unsigned log_r,
                        unsigned coset_first, unsigned coset_step, unsigned num_cosets, size_t ncols, u64 *d_out, u64 *d_tmp) {
  int la, lb;
  if (!ntt_tile_supported((int)log_n, &la, &lb)) return false;
  const size_t n = (size_t)1 << log_n, n2 = (size_t)1 << la, n1 = (size_t)1 << lb;
  const int ra = tile_logr(la), rb = tile_logr(lb);
  if (fold_count > (n1 << ra) || fold_count > n) return false;
  const unsigned TA = TILE_THREADS >> ra, TB = TILE_THREADS >> rb;
  if (n1 % TA || n2 % TB) return false;
  const u64 wrn = root_of_unity_mont(log_n + log_r);
  PowTab pre = c.get_pow_tab(wrn, (int)(log_n + log_r));
  const u64 mask = ((u64)1 << (log_n + log_r)) - 1;
  // prescale table + fold factors of this coset set (cached per job shape)
  u64 *S = c.get_prescale(log_n, log_r, (unsigned)lb, (unsigned)n2, coset_first, coset_step, num_cosets, [&](u64 *dS, u64 *dff) {
    prescale_table_kernel<<<dim3((unsigned)((n2 + 255) / 256), num_cosets), 256, 0, c.stream>>>(dS, dff, pre, log_n, (unsigned)lb,
                                                                                                  (unsigned)n2, coset_first, coset_step, mask);
    c.launches++;
    TVM_CUDA(cudaGetLastError());
  });
  u64 *ff = S + (size_t)num_cosets * n2;
  TileJob a{};
  a.in = d_coef; a.out = d_tmp;
  a.in_col_stride = coef_stride; a.in_coset_stride = 0;
  a.out_col_stride = (size_t)num_cosets * n; a.out_coset_stride = n;
  a.in_row_stride = 1; a.in_elem_stride = n1;
  a.out_row_stride = 1; a.out_elem_stride = n1;
  a.tw = c.get_tile_tw(la, false);
  a.prescale = S; a.fold_factor = ff; a.fold_count = fold_count; a.fold_offset = n;
  a.post_mode = 1; a.G = pre; a.exp_mask = mask; a.mul = (u64)1 << log_r;
  a.coset_first = coset_first; a.coset_step = coset_step;
  TileJob b{};
  b.in = d_tmp; b.out = d_out;
  b.in_col_stride = (size_t)num_cosets * n; b.in_coset_stride = n;
  b.out_col_stride = (size_t)num_cosets * n; b.out_coset_stride = n;
  b.in_row_stride = n1; b.in_elem_stride = 1;
  b.out_row_stride = 1; b.out_elem_stride = n2;
  b.tw = c.get_tile_tw(lb, false);
  b.post_mode = 0;
  // launch pairs of (columns x cosets) whose transposed intermediate fits the budget
  const size_t budget_words = tile_tmp_budget() / 8;
  size_t cos_per = std::min<size_t>(num_cosets, std::max<size_t>(1, budget_words / n));
  while (num_cosets % cos_per) cos_per--;
  const size_t group = cos_per == num_cosets ? std::max<size_t>(1, budget_words / ((size_t)num_cosets * n)) : 1;
  const size_t pair_words = group * cos_per * n;
  const size_t npairs = ((ncols + group - 1) / group) * (num_cosets / cos_per);
  const bool two = tile_streams() == 2 && npairs >= 2 && 2 * pair_words <= ncols * (size_t)num_cosets * n;
  cudaStream_t st[2] = {c.stream, c.stream};
  if (two) {
    st[1] = c.get_pair_stream();
    TVM_CUDA(cudaEventRecord(c.pair_fork, c.stream));
    TVM_CUDA(cudaStreamWaitEvent(st[1], c.pair_fork, 0));
  }
  size_t pi = 0;
  for (size_t c0 = 0; c0 < ncols; c0 += group) {
    const size_t g = std::min(group, ncols - c0);
    for (size_t y0 = 0; y0 < num_cosets; y0 += cos_per, pi++) {
      const int k = two ? (int)(pi & 1) : 0;
      TileJob aa = a, bb = b;
      aa.in += c0 * coef_stride;
      aa.out += (size_t)k * pair_words;
      aa.prescale += y0 * n2; aa.fold_factor += y0;
      aa.coset_first = coset_first + coset_step * (unsigned)y0;
      aa.out_col_stride = cos_per * n;                       // the intermediate of this pair only
      bb.in += (size_t)k * pair_words;
      bb.in_col_stride = cos_per * n;
      bb.out += (c0 * (size_t)num_cosets + y0) * n;
      launch_tile<false, true, 1, false>(ra, dim3((unsigned)(n1 / TA), (unsigned)cos_per, (unsigned)g), st[k], aa);
      launch_tile<false, false, 0, true>(rb, dim3((unsigned)(n2 / TB), (unsigned)cos_per, (unsigned)g), st[k], bb);
      c.launches += 2;
    }
  }
  if (two) {
    TVM_CUDA(cudaEventRecord(c.pair_join, st[1]));
    TVM_CUDA(cudaStreamWaitEvent(c.stream, c.pair_join, 0));
  }
  return true;
}

// Interpolation of trace columns to pre-scaled coefficients (lde_interpolate_run's contract).
bool lde_interpolate_tiles(Ctx &c, const u64 *d_trace, const u64 *d_rand, unsigned num_rand, unsigned rand_pad, unsigned log_n,
                           u64 offset_mont, size_t ncols, u64 *d_coef, size_t coef_stride, u64 *d_tmp) {
  int la, lb;
  if (!ntt_tile_supported((int)log_n, &la, &lb)) return false;
  const size_t n = (size_t)1 << log_n, n2 = (size_t)1 << la, n1 = (size_t)1 << lb;
  const int ra = tile_logr(la), rb = tile_logr(lb);
  const unsigned TA = TILE_THREADS >> ra, TB = TILE_THREADS >> rb;
  if (n1 % TA || n2 % TB) return false;
  const u64 wn_inv = finv(root_of_unity_mont(log_n));
  TileJob a{};
  a.in = d_trace; a.out = d_tmp;
  a.in_col_stride = n; a.out_col_stride = n;
  a.in_row_stride = 1; a.in_elem_stride = n1;
  a.out_row_stride = 1; a.out_elem_stride = n1;
  a.tw = c.get_tile_tw(la, true);
  a.post_mode = 1; a.G = c.get_pow_tab(wn_inv, (int)log_n); a.exp_mask = n - 1; a.mul = 1;
  TileJob b{};
  b.in = d_tmp; b.out = d_coef;
  b.in_col_stride = n; b.out_col_stride = coef_stride;
  b.in_row_stride = n1; b.in_elem_stride = 1;
  b.out_row_stride = 1; b.out_elem_stride = n2;
  b.tw = c.get_tile_tw(lb, true);
  b.post_mode = 2; b.G = c.get_pow_tab(offset_mont, (int)log_n + 1); b.mul = n2;
  b.scalar = finv(to_mont((u64)n));
  size_t group = std::max<size_t>(1, tile_tmp_budget() / 8 / n);
  for (size_t c0 = 0; c0 < ncols; c0 += group) {
    const size_t g = std::min(group, ncols - c0);
    TileJob aa = a, bb = b;
    aa.in += c0 * n;
    bb.out += c0 * coef_stride;
    launch_tile<true, false, 1, false>(ra, dim3((unsigned)(n1 / TA), 1, (unsigned)g), c.stream, aa);
    launch_tile<true, false, 2, false>(rb, dim3((unsigned)(n2 / TB), 1, (unsigned)g), c.stream, bb);
    c.launches += 2;
  }
  const unsigned pad = d_rand ? std::max(rand_pad, num_rand) : 0;
  if (pad) {
    lde_randomizer_kernel<<<dim3((pad + 127) / 128, (unsigned)ncols), 128, 0, c.stream>>>(d_coef, coef_stride, n, d_rand, num_rand, pad, b.G);
    c.launches++;
    TVM_CUDA(cudaGetLastError());
  }
  return true;
}

}  // namespace tvm

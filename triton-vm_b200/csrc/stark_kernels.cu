// Streaming (HBM-bound) kernels of the prover between the big transforms:
// quotient-segment randomisation, column dot products (out-of-domain rows, polynomial
// evaluation), weighted column sums, DEEP combination, FRI folding, row gathering.
// Each cites the reference function it replaces; all data Montgomery form, X-field vectors are
// planar (coordinate d of element i at base + d*stride + i).
#include "ctx.h"
#include "launch.h"
#include "stark.h"

namespace tvm {

static constexpr int EW_THREADS = 256;
static inline unsigned ew_grid(size_t n) { return (unsigned)((n + EW_THREADS - 1) / EW_THREADS); }

__device__ __forceinline__ u64 pt(const PowTab &t, u64 e) {
  return fmul(__ldg(t.lo + (e & ((1ULL << t.shift) - 1))), __ldg(t.hi + (e >> t.shift)));
}

// ---- layout ------------------------------------------------------------------------------
// Sharded coset-major data -> natural order.  Rank g of `world` owns the domain cosets c with
// c mod world == g, stored locally as [y][k], y = c div world; after an all-gather the buffer is
// [rank][...local block...].  Natural index i = c + r*k.
//   planar variant: element (rank, plane d, y, k) at in[rank*rank_stride + d*plane_stride + y*n + k]
__global__ void shards_to_natural_kernel(const u64 *in, u64 *out, size_t rank_stride, size_t plane_stride, size_t out_stride, int log_n,
                                         int log_r, int log_w, int planes) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)1 << (log_n + log_r);
  if (i >= total) return;
  size_t c = i & (((size_t)1 << log_r) - 1), k = i >> log_r;
  size_t rank = c & (((size_t)1 << log_w) - 1), y = c >> log_w;
  const u64 *src = in + rank * rank_stride + (y << log_n) + k;
  for (int d = 0; d < planes; d++) out[d * out_stride + i] = src[d * plane_stride];
}
void shards_to_natural_run(Ctx &c, const u64 *in, u64 *out, size_t rank_stride, size_t plane_stride, size_t out_stride, int log_n, int log_r,
                           int log_w, int planes) {
  shards_to_natural_kernel<<<ew_grid((size_t)1 << (log_n + log_r)), EW_THREADS, 0, c.stream>>>(in, out, rank_stride, plane_stride,
                                                                                                  out_stride, log_n, log_r, log_w, planes);
  c.launches++;
  TVM_CUDA(cudaGetLastError());
}
void coset_to_natural_run(Ctx &c, const u64 *in, u64 *out, size_t in_stride, size_t out_stride, int log_n, int log_r, int planes) {
  shards_to_natural_run(c, in, out, 0, in_stride, out_stride, log_n, log_r, 0, planes);
}
//   record variant (row digests): record of row (rank, y, k) at in[(rank*rows_per_rank + y*n + k)*5 .. +5]
__global__ void shard_digests_to_natural_kernel(const u64 *in, u64 *out, int log_n, int log_r, int log_w) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)5 << (log_n + log_r);
  if (t >= total) return;
  size_t i = t / 5, e = t - 5 * i;
  size_t c = i & (((size_t)1 << log_r) - 1), k = i >> log_r;
  size_t rank = c & (((size_t)1 << log_w) - 1), y = c >> log_w;
  size_t rows_per_rank = (size_t)1 << (log_n + log_r - log_w);
  out[t] = in[(rank * rows_per_rank + (y << log_n) + k) * 5 + e];
}
void shard_digests_to_natural_run(Ctx &c, const u64 *in, u64 *out, int log_n, int log_r, int log_w) {
  shard_digests_to_natural_kernel<<<ew_grid((size_t)5 << (log_n + log_r)), EW_THREADS, 0, c.stream>>>(in, out, log_n, log_r, log_w);
  c.launches++;
  TVM_CUDA(cudaGetLastError());
}

// [rows][3] interleaved X-field (ABI layout) <-> 3 planes; used for aux trace / randomizer upload
__global__ void deinterleave3_kernel(const u64 *in, u64 *out, size_t len, size_t ncols) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over ncols*len elements
  if (idx >= ncols * len) return;
  size_t q = idx / len, j = idx - q * len;
  const u64 *src = in + 3 * idx;
  for (int d = 0; d < 3; d++) out[(3 * q + d) * len + j] = src[d];
}
void deinterleave3_run(Ctx &c, const u64 *in, u64 *out, size_t len, size_t ncols) {
  deinterleave3_kernel<<<ew_grid(ncols * len), EW_THREADS, 0, c.stream>>>(in, out, len, ncols);
  c.launches++;
  TVM_CUDA(cudaGetLastError());
}

// ---- quotient segments ----------------------------------------------------------------------
// Input: quotient interpolant coefficients a_j (planar, natural order, length 4*seg_len, already
// divided by offset^j, i.e. true coefficients).  Output: the 5 randomized segment polynomials
// s_0..s_4 (stark.rs:1252-1263 split, 1302-1356 randomisation), each PRE-SCALED by offset^j so
// they can feed the coset-LDE directly: out[(3*s+d)*out_stride + j] = s_s[j].coord_d * offset^j.
//   s_4 = randomizer;  s_i = q_i - zeta^i * s_{i+1}(zeta^4 X),  q_i[j] = a[i + 4j].
__global__ void segment_chain_kernel(SegmentArgs a) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.seg_len) return;
  u64 z4j = pt(a.zeta4, j), oj = pt(a.off, j);
  xfe nxt = xzero();
  if (j < a.rnd_len) nxt = xmake(a.rnd[j], a.rnd[a.rnd_stride + j], a.rnd[2 * a.rnd_stride + j]);
  u64 *o = a.out + j;
  xfe s4 = xmulb(nxt, oj);
  o[12 * a.out_stride] = s4.c0; o[13 * a.out_stride] = s4.c1; o[14 * a.out_stride] = s4.c2;
#pragma unroll
  for (int i = 3; i >= 0; i--) {
    size_t src = (size_t)i + 4 * j;
    xfe q = xmake(a.quot[src], a.quot[a.quot_stride + src], a.quot[2 * a.quot_stride + src]);
    xfe cur = xsub(q, xmulb(nxt, fmul(a.zeta_pow[i], z4j)));
    xfe sc = xmulb(cur, oj);
    o[(3 * i) * a.out_stride] = sc.c0; o[(3 * i + 1) * a.out_stride] = sc.c1; o[(3 * i + 2) * a.out_stride] = sc.c2;
    nxt = cur;
  }
}
void segment_chain_run(Ctx &c, const SegmentArgs &a) {
  segment_chain_kernel<<<ew_grid(a.seg_len), EW_THREADS, 0, c.stream>>>(a);
  c.launches++;
  TVM_CUDA(cudaGetLastError());
}

// ---- power vector: out[j] = base^j (X-field, planar) ------------------------------------------
__global__ void xpow_vector_kernel(xfe base, u64 *out, size_t stride, size_t len) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= len) return;
  xfe r = xpow(base, j);
  out[j] = r.c0; out[stride + j] = r.c1; out[2 * stride + j] = r.c2;
}
void xpow_vector_run(Ctx &c, xfe base, u64 *out, size_t stride, size_t len) {
  xpow_vector_kernel<<<ew_grid(len), EW_THREADS, 0, c.stream>>>(base, out, stride, len);
  c.launches++;
  TVM_CUDA(cudaGetLastError());
}

// ---- column dot products: out[q] = sum_j col_q[j] * xvec[j]  (B-field columns x X-field vector)
// Replaces the batched barycentric evaluation of MasterTable::out_of_domain_row
// (master_table.rs:348-390) and Polynomial::evaluate on segment / combination polynomials
// (stark.rs:474-494, 566-607): with c'_j = c_j*offset^j stored, f(beta) = sum c'_j (beta/offset)^j.
// grid (ncols, split): each CTA covers a slice of one column for all `nvec` vectors; products are
// accumulated unreduced (128 bits + overflow word per coordinate) and reduced once per thread;
// a second kernel adds the `split` partial results of a column.
static constexpr int DOT_THREADS = 256;
__device__ __forceinline__ void dot_mac(u64 &lo, u64 &hi, u32 &ov, u64 x, u64 y) {
  const unsigned __int128 p128 = (unsigned __int128)x * y;   // one 128-bit product: 4 IMAD.WIDE (x * y and __umul64hi form the low half twice)
  u64 plo = (u64)p128, phi = (u64)(p128 >> 64);
  asm("add.cc.u64 %0, %0, %3;\n\taddc.cc.u64 %1, %1, %4;\n\taddc.u32 %2, %2, 0;"
      : "+l"(lo), "+l"(hi), "+r"(ov) : "l"(plo), "l"(phi));
}
__device__ __forceinline__ u64 dot_reduce160(u64 lo, u64 hi, u32 ov) {
  // (lo + 2^64 hi + 2^128 ov) * 2^-64 = montyred(lo, 0) + hi + ov * 2^64  (mod p)
  u64 h = hi >= P ? hi - P : hi;
  return fadd(fadd(montyred(lo, 0), h), (u64)ov * EPS);
}
template <int NVEC>
__global__ void __launch_bounds__(DOT_THREADS) col_dot_kernel(const u64 *cols, size_t col_stride, size_t len, const u64 *xvec,
                                                              size_t xvec_stride, size_t xvec_set_stride, u64 *partial) {
  const u64 *col = cols + (size_t)blockIdx.x * col_stride;
  const size_t per = (len + gridDim.y - 1) / gridDim.y;
  const size_t j0 = (size_t)blockIdx.y * per, j1 = min(len, j0 + per);
  u64 lo[NVEC][3], hi[NVEC][3];
  u32 ov[NVEC][3];
#pragma unroll
  for (int v = 0; v < NVEC; v++)
#pragma unroll
    for (int d = 0; d < 3; d++) { lo[v][d] = 0; hi[v][d] = 0; ov[v][d] = 0; }
  for (size_t j = j0 + threadIdx.x; j < j1; j += blockDim.x) {
    u64 cv = col[j];
#pragma unroll
    for (int v = 0; v < NVEC; v++) {
      const u64 *xv = xvec + v * xvec_set_stride + j;
#pragma unroll
      for (int d = 0; d < 3; d++) dot_mac(lo[v][d], hi[v][d], ov[v][d], xv[d * xvec_stride], cv);
    }
  }
  __shared__ u64 red[DOT_THREADS * 3];
#pragma unroll
  for (int v = 0; v < NVEC; v++) {
#pragma unroll
    for (int d = 0; d < 3; d++) red[d * DOT_THREADS + threadIdx.x] = dot_reduce160(lo[v][d], hi[v][d], ov[v][d]);
    __syncthreads();
    for (int s = DOT_THREADS / 2; s > 0; s >>= 1) {
      if ((int)threadIdx.x < s) {
#pragma unroll
        for (int d = 0; d < 3; d++)
          red[d * DOT_THREADS + threadIdx.x] = fadd(red[d * DOT_THREADS + threadIdx.x], red[d * DOT_THREADS + threadIdx.x + s]);
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      u64 *o = partial + (((size_t)blockIdx.x * gridDim.y + blockIdx.y) * NVEC + v) * 3;
      o[0] = red[0]; o[1] = red[DOT_THREADS]; o[2] = red[2 * DOT_THREADS];
    }
    __syncthreads();
  }
}
// out[i] = sum_s partial[(i / w) * split * w + s * w + i % w],  w = nvec*3 words per column
__global__ void col_dot_finish_kernel(const u64 *partial, unsigned split, unsigned w, size_t total, u64 *out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  size_t col = i / w, r = i - col * w;
  u64 acc = 0;
  for (unsigned s = 0; s < split; s++) acc = fadd(acc, partial[(col * split + s) * w + r]);
  out[i] = acc;
}
// out: [ncols][nvec][3]
void col_dot_run(Ctx &c, const u64 *cols, size_t col_stride, size_t ncols, size_t len, const u64 *xvec, size_t xvec_stride,
                 size_t xvec_set_stride, int nvec, u64 *out) {
  if (!ncols) return;
  if (nvec != 1 && nvec != 2) throw ApiError{TVM_ERR_INVALID_ARG, "col_dot: nvec must be 1 or 2"};
  // enough CTAs for ~8 per SM, but at least 4 elements per thread
  size_t split = std::max<size_t>(1, std::min<size_t>((148 * 8 + ncols - 1) / ncols, len / (4 * DOT_THREADS) + 1));
  u64 *partial = (u64 *)c.pool_alloc(sizeof(u64) * ncols * split * nvec * 3);   // stream-ordered reuse: released below
  dim3 grid((unsigned)ncols, (unsigned)split);
  if (nvec == 1) col_dot_kernel<1><<<grid, DOT_THREADS, 0, c.stream>>>(cols, col_stride, len, xvec, xvec_stride, xvec_set_stride, partial);
  else col_dot_kernel<2><<<grid, DOT_THREADS, 0, c.stream>>>(cols, col_stride, len, xvec, xvec_stride, xvec_set_stride, partial);
  size_t total = ncols * nvec * 3;
  col_dot_finish_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c.stream>>>(partial, (unsigned)split, (unsigned)(nvec * 3), total, out);
  c.launches += 2;
  c.pool_release(partial);
  TVM_CUDA(cudaGetLastError());
}

// ---- weighted column sum: out[j] = sum_q w_q * col_q[j] ------------------------------------------
// MasterTable::weighted_sum_of_columns (master_table.rs:512-542) on the stored (pre-scaled)
// interpolant coefficients; B-field columns (main) and X-field columns given as 3 planar
// B-field columns (aux).  Accumulates into out when `accumulate`.
__global__ void weighted_colsum_kernel(const u64 *cols, size_t col_stride, unsigned ncols, int xfield, const u64 *w /*[ncols][3]*/,
                                       size_t len, u64 *out, size_t out_stride, int accumulate) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= len) return;
  xfe acc = xzero();
  if (accumulate) acc = xmake(out[j], out[out_stride + j], out[2 * out_stride + j]);
  for (unsigned q = 0; q < ncols; q++) {
    xfe wq = xmake(__ldg(w + 3 * q), __ldg(w + 3 * q + 1), __ldg(w + 3 * q + 2));
    if (xfield) {
      const u64 *p = cols + (size_t)(3 * q) * col_stride + j;
      acc = xadd(acc, xmul(wq, xmake(p[0], p[col_stride], p[2 * col_stride])));
    } else {
      acc = xadd(acc, xmulb(wq, cols[(size_t)q * col_stride + j]));
    }
  }
  out[j] = acc.c0; out[out_stride + j] = acc.c1; out[2 * out_stride + j] = acc.c2;
}
void weighted_colsum_run(Ctx &c, const u64 *cols, size_t col_stride, unsigned ncols, bool xfield, const u64 *d_w, size_t len, u64 *out,
                         size_t out_stride, bool accumulate) {
  weighted_colsum_kernel<<<ew_grid(len), EW_THREADS, 0, c.stream>>>(cols, col_stride, ncols, xfield, d_w, len, out, out_stride,
                                                                     accumulate);
  c.launches++;
  TVM_CUDA(cudaGetLastError());
}

// ---- DEEP combination (stark.rs:545-639, 1360-1379, 2096-2103) --------------------------------------
// in: 3 X-field codewords (main&aux combination, p, r) as 9 coset-major planes holding this rank's
// cosets; out: the combination codeword in the same (local coset-major, planar) order.
__global__ void deep_kernel(DeepArgs a) {
  size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)a.num_cosets << a.log_n;
  if (m >= total) return;
  size_t y = m >> a.log_n, k = m & (((size_t)1 << a.log_n) - 1);
  size_t i = (a.coset_first + a.coset_step * y) + (k << a.log_r);
  u64 x = fmul(a.offset, pt(a.dom, i));
  const u64 *p = a.cw + m;
  xfe f[3];
#pragma unroll
  for (int v = 0; v < 3; v++) f[v] = xmake(p[(3 * v) * a.cw_stride], p[(3 * v + 1) * a.cw_stride], p[(3 * v + 2) * a.cw_stride]);
  const int src[4] = {0, 0, 1, 2};
  // one X-field inversion for the four denominators (Montgomery's trick)
  xfe den[4], pre[4];
  xfe run = xone();
#pragma unroll
  for (int t = 0; t < 4; t++) {
    den[t] = xneg(a.point[t]);
    den[t].c0 = fadd(den[t].c0, x);                 // x - point
    pre[t] = run;
    run = xmul(run, den[t]);
  }
  xfe inv = xinv(run);
  xfe acc = xzero();
#pragma unroll
  for (int t = 3; t >= 0; t--) {
    xfe dinv = xmul(inv, pre[t]);
    inv = xmul(inv, den[t]);
    xfe term = xmul(xsub(f[src[t]], a.value[t]), dinv);
    acc = xadd(acc, xmul(term, a.weight[t]));
  }
  a.out[m] = acc.c0; a.out[a.out_stride + m] = acc.c1; a.out[2 * a.out_stride + m] = acc.c2;
}
void deep_run(Ctx &c, const DeepArgs &a) {
  deep_kernel<<<ew_grid((size_t)a.num_cosets << a.log_n), EW_THREADS, 0, c.stream>>>(a);
  c.launches++;
  TVM_CUDA(cudaGetLastError());
}

// ---- FRI (fri.rs:343-366) ------------------------------------------------------------------------------
// leaves: Digest::from(xfe) = (c0,c1,c2,0,0) written straight into the node array
__global__ void fri_leaves_kernel(const u64 *cw, size_t stride, size_t n, u64 *leaves) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u64 *d = leaves + 5 * i;
  d[0] = cw[i]; d[1] = cw[stride + i]; d[2] = cw[2 * stride + i]; d[3] = 0; d[4] = 0;
}
void fri_leaves_run(Ctx &c, const u64 *cw, size_t stride, size_t n, u64 *leaves) {
  fri_leaves_kernel<<<ew_grid(n), EW_THREADS, 0, c.stream>>>(cw, stride, n, leaves);
  c.launches++;
  TVM_CUDA(cudaGetLastError());
}
// split_and_fold: out[i] = ((1 + c/x_i) f[i] + (1 - c/x_i) f[i + n/2]) / 2,  x_i = offset * g^i
__global__ void fri_fold_kernel(const u64 *in, size_t in_stride, size_t n, u64 offset_inv, PowTab ginv, xfe chal, u64 two_inv, u64 *out,
                                size_t out_stride) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t half = n >> 1;
  if (i >= half) return;
  u64 xinv_i = fmul(offset_inv, pt(ginv, i));
  xfe s = xmulb(chal, xinv_i);
  xfe a = xmake(in[i], in[in_stride + i], in[2 * in_stride + i]);
  xfe b = xmake(in[half + i], in[in_stride + half + i], in[2 * in_stride + half + i]);
  xfe l = xmul(xaddb(s, MONT_ONE), a);
  xfe r = xmul(xaddb(xneg(s), MONT_ONE), b);
  xfe o = xmulb(xadd(l, r), two_inv);
  out[i] = o.c0; out[out_stride + i] = o.c1; out[2 * out_stride + i] = o.c2;
}
void fri_fold_run(Ctx &c, const u64 *in, size_t in_stride, size_t n, u64 offset_mont, xfe chal, u64 *out, size_t out_stride) {
  int log_n = 0;
  while (((size_t)1 << log_n) < n) log_n++;
  PowTab ginv = c.get_pow_tab(finv(root_of_unity_mont(log_n)), log_n);
  u64 two_inv = finv(fadd(MONT_ONE, MONT_ONE));
  fri_fold_kernel<<<ew_grid(n / 2), EW_THREADS, 0, c.stream>>>(in, in_stride, n, finv(offset_mont), ginv, chal, two_inv, out, out_stride);
  c.launches++;
  TVM_CUDA(cudaGetLastError());
}

// ---- gathers (MasterTable::reveal_rows, master_table.rs:548-555; authentication structures) ----
// rows[t][q] = table[q][mem_index(idx[t])], canonical output
// Sharded tables (log_w > 0): a row this rank does not own is written as zeros, so that the
// wrapping sum over ranks (all_reduce_sum_u64) reassembles the rows exactly.
__global__ void gather_rows_kernel(const u64 *table, size_t col_stride, unsigned ncols, const unsigned *idx, unsigned nidx, int log_n,
                                   int log_r, int log_w, unsigned rank, unsigned coset_mem_stride, u64 *out) {
  size_t t = blockIdx.x;
  if (t >= nidx) return;
  size_t i = idx[t];
  size_t m = i;
  bool mine = true;
  if (log_r >= 0) {
    size_t c = i & (((size_t)1 << log_r) - 1), k = i >> log_r;
    mine = (c & (((size_t)1 << log_w) - 1)) == rank;
    m = (((c >> log_w) * coset_mem_stride) << log_n) + k;
  }
  for (unsigned q = threadIdx.x; q < ncols; q += blockDim.x)
    out[t * ncols + q] = mine ? from_mont(table[(size_t)q * col_stride + m]) : 0;
}
void gather_rows_run(Ctx &c, const u64 *table, size_t col_stride, unsigned ncols, const unsigned *d_idx, unsigned nidx, int log_n, int log_r,
                     u64 *d_out, int log_w, unsigned rank, unsigned coset_mem_stride) {
  if (!nidx) return;
  gather_rows_kernel<<<nidx, 128, 0, c.stream>>>(table, col_stride, ncols, d_idx, nidx, log_n, log_r, log_w, rank, coset_mem_stride, d_out);
  c.launches++;
  TVM_CUDA(cudaGetLastError());
}
// Sharded trees (log_w > 0, merkle_run_sharded): a node of a level at least 2^log_w wide belongs to the rank whose
// subtree contains it; the replicated top levels are contributed by rank 0 only; everything else is written as zeros
// so that all_reduce_sum_u64 reassembles the digests.
// low-memory mode: rows of ONE coset table [ncols][n]; kt = (k, t) pairs: out[t][q] = table[q][k]
__global__ void gather_rows_scatter_kernel(const u64 *table, size_t col_stride, unsigned ncols, const unsigned *kt, unsigned count, u64 *out) {
  unsigned j = blockIdx.x;
  if (j >= count) return;
  const size_t k = kt[2 * j], t = kt[2 * j + 1];
  for (unsigned q = threadIdx.x; q < ncols; q += blockDim.x) out[t * ncols + q] = from_mont(table[(size_t)q * col_stride + k]);
}
void gather_rows_scatter_run(Ctx &c, const u64 *table, size_t col_stride, unsigned ncols, const unsigned *d_kt, unsigned count, u64 *d_out) {
  if (!count) return;
  gather_rows_scatter_kernel<<<count, 128, 0, c.stream>>>(table, col_stride, ncols, d_kt, count, d_out);
  c.launches++;
  TVM_CUDA(cudaGetLastError());
}
__global__ void gather_digests_kernel(const u64 *nodes, const unsigned *idx, unsigned nidx, u64 *out, int log_w, unsigned rank) {
  unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nidx * 5) return;
  const unsigned i = idx[t / 5];
  bool mine = true;
  if (log_w > 0) {
    const int level = 31 - __clz(i);                       // node i is at level floor(log2 i), width 2^level
    if (level >= log_w) mine = ((i - (1u << level)) >> (level - log_w)) == rank;
    else mine = rank == 0;
  }
  out[t] = mine ? from_mont(nodes[(size_t)i * 5 + t % 5]) : 0;
}
void gather_digests_run(Ctx &c, const u64 *nodes, const unsigned *d_idx, unsigned nidx, u64 *d_out, int log_w, unsigned rank) {
  if (!nidx) return;
  gather_digests_kernel<<<ew_grid((size_t)nidx * 5), EW_THREADS, 0, c.stream>>>(nodes, d_idx, nidx, d_out, log_w, rank);
  c.launches++;
  TVM_CUDA(cudaGetLastError());
}

// planar vector scale: v[j] *= s^j-independent scalar powers (used to undo the coset offset after iNTT):
// out[d][j] = in[d][j] * tab^j
__global__ void scale_by_powers_kernel(u64 *v, size_t stride, int planes, size_t len, PowTab tab) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= len) return;
  u64 s = pt(tab, j);
  for (int d = 0; d < planes; d++) v[d * stride + j] = fmul(v[d * stride + j], s);
}
void scale_by_powers_run(Ctx &c, u64 *v, size_t stride, int planes, size_t len, PowTab tab) {
  scale_by_powers_kernel<<<ew_grid(len), EW_THREADS, 0, c.stream>>>(v, stride, planes, len, tab);
  c.launches++;
  TVM_CUDA(cudaGetLastError());
}

}  // namespace tvm

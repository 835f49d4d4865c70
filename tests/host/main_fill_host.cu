// TEST INFRASTRUCTURE — not part of the library.
// The table-fill stages of triton-vm_b200/csrc/fill/main_fill.cuh instantiated with a sequential host executor, so that
// `pytest -m "not gpu"` can check their logic (sorting order, padding rules, u32 sections, the Bezout polynomial passes)
// against the oracle's table fill on a machine without a GPU.  The library itself (csrc/main_fill.cu) instantiates the same
// header with the CUDA executor only; nothing here is linked into libtvm_b200.so.
// Build: nvcc -O2 -std=c++17 --extended-lambda -shared -Xcompiler -fPIC tests/host/main_fill_host.cu -o tests/host/_build/libmain_fill_host.so
#include <algorithm>
#include <cstring>
#include <memory>
#include <numeric>
#include <vector>
#include "../../include/tvm_b200.h"
#include "../../triton-vm_b200/csrc/fill/main_fill.cuh"
#include "../../triton-vm_b200/csrc/tip5_constants.inc"

using namespace tvm;

namespace {

struct HostExec {
  std::vector<std::unique_ptr<u64[]>> mem;
  size_t launches = 0, transforms = 0;
  u64 *alloc(size_t words) {
    mem.emplace_back(new u64[words ? words : 1]);
    std::memset(mem.back().get(), 0xA5, (words ? words : 1) * 8);   // uninitialised reads show up as garbage
    return mem.back().get();
  }
  template <class F>
  void launch(size_t count, F f) {
    launches++;
    for (size_t i = 0; i < count; i++) f(i);
  }
  void sort_perm(const u64 *keys, u64 *perm, size_t n) {
    std::vector<u64> idx(n);
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](u64 a, u64 b) { return keys[a] < keys[b]; });
    std::copy(idx.begin(), idx.end(), perm);
  }
  void exclusive_sum(const u64 *in, u64 *out, size_t n) {
    u64 acc = 0;
    for (size_t i = 0; i < n; i++) { const u64 v = in[i]; out[i] = acc; acc += v; }
  }
  // radix-2 transform with w = 7^((p-1)/N): any primitive root serves the products and correlations of the fill
  void ntt(const u64 *in, u64 *out, unsigned lg, size_t count, bool inverse) {
    transforms += count;
    const size_t N = (size_t)1 << lg;
    u64 w = fpow(to_mont(7), (P - 1) >> lg);
    if (inverse) w = finv(w);
    const u64 ninv = finv(to_mont((u64)N));
    for (size_t c = 0; c < count; c++) {
      const u64 *a = in + c * N;
      u64 *o = out + c * N;
      for (size_t i = 0; i < N; i++) {
        size_t r = 0;
        for (unsigned b = 0; b < lg; b++) r |= ((i >> b) & 1) << (lg - 1 - b);
        o[r] = a[i];
      }
      for (size_t len = 2; len <= N; len <<= 1) {
        const u64 wl = fpow(w, N / len);
        for (size_t s = 0; s < N; s += len) {
          u64 t = MONT_ONE;
          for (size_t j = 0; j < len / 2; j++) {
            const u64 u = o[s + j], v = fmul(o[s + j + len / 2], t);
            o[s + j] = fadd(u, v);
            o[s + j + len / 2] = fsub(u, v);
            t = fmul(t, wl);
          }
        }
      }
      if (inverse) for (size_t i = 0; i < N; i++) o[i] = fmul(o[i], ninv);
    }
  }
  u64 read_word(const u64 *p) { return *p; }
};

std::vector<u64> constants() {
  static const u64 rc_mont[80] = {TVM_TIP5_RC_MONT};
  static const unsigned char lut[256] = {TVM_TIP5_LUT};
  u64 rc0[16];
  for (int k = 0; k < 16; k++) rc0[k] = from_mont(rc_mont[k]);
  return fill::fill_constants(rc0, lut);
}

}  // namespace

extern "C" {

// [149][n] canonical table columns of the AET; returns 0, or 1 with `err` filled
int fill_host_main_table(const tvm_aet *a, uint64_t n, uint64_t *out, uint64_t *lengths9, unsigned bezout_direct_log, char *err, size_t err_cap) {
  try {
    HostExec ex;
    std::vector<u64> mult(a->program_len);
    for (size_t i = 0; i < a->program_len; i++) mult[i] = a->instruction_multiplicities[i];
    fill::AetView v{a->program, (size_t)a->program_len, mult.data(), a->processor_trace, (size_t)a->processor_rows,
                    a->op_stack_underflow_trace, (size_t)a->op_stack_rows, a->ram_trace, (size_t)a->ram_rows,
                    a->program_hash_trace, (size_t)a->program_hash_rows, a->sponge_trace, (size_t)a->sponge_rows,
                    a->hash_trace, (size_t)a->hash_rows, a->u32_entries, (size_t)a->u32_count,
                    a->cascade_table_lookup_multiplicities, (size_t)a->cascade_count, a->lookup_table_lookup_multiplicities};
    const std::vector<u64> consts = constants();
    const fill::FillInfo info = fill::main_table_from_aet(ex, v, consts.data(), (size_t)n, out, bezout_direct_log);
    if (lengths9) {
      const uint64_t l[9] = {info.program_len_padded, info.processor_len, info.op_stack_len, info.ram_len, info.processor_len,
                             info.hash_len, info.cascade_len, 256, info.u32_len};
      std::memcpy(lengths9, l, sizeof l);
    }
    return 0;
  } catch (const std::exception &e) {
    if (err && err_cap) { std::strncpy(err, e.what(), err_cap - 1); err[err_cap - 1] = 0; }
    return 1;
  }
}

// canonical in, canonical out; stats (optional) [2]: bodies launched, transforms
int fill_host_bezout(const uint64_t *roots, uint64_t m, uint64_t *a_out, uint64_t *b_out, unsigned direct_log, uint64_t *stats) {
  try {
    HostExec ex;
    std::vector<u64> r(m), a(m), b(m);
    for (size_t i = 0; i < m; i++) r[i] = to_mont(roots[i]);
    fill::bezout_coefficients(ex, r.data(), (size_t)m, a.data(), b.data(), direct_log);
    for (size_t i = 0; i < m; i++) { a_out[i] = from_mont(a[i]); b_out[i] = from_mont(b[i]); }
    if (stats) { stats[0] = ex.launches; stats[1] = ex.transforms; }
    return 0;
  } catch (const std::exception &) {
    return 1;
  }
}

}  // extern "C"

// Register-resident small DFTs over GF(p), p = 2^64 - 2^32 + 1, whose twiddles are all powers of two.
//
// 2 has multiplicative order 192 mod p (2^96 = -1), so every 64th root of unity is a power of two; with
// twenty-first's PRIMITIVE_ROOTS (the roots ntt/intt use, reference call sites stark.rs:872-877):
//     w_2 = 2^96, w_4 = 2^48, w_8 = 2^120, w_16 = 2^156, w_32 = 2^78, w_64 = 2^39.
// A 2^r-point DFT (r <= 6) therefore needs no multiplier at all: radix-2 DIF stages whose twiddles are
// shift-and-reduce (fmul_2k).  The code is TVM_HD so that tests can run it on the host.
#pragma once
#include "field.cuh"

namespace tvm {

// exponent e with w_{2^r} = 2^e (mod p)
template <int R> struct RootLog2;
template <> struct RootLog2<1> { static constexpr int E = 96; };
template <> struct RootLog2<2> { static constexpr int E = 48; };
template <> struct RootLog2<3> { static constexpr int E = 120; };
template <> struct RootLog2<4> { static constexpr int E = 156; };
template <> struct RootLog2<5> { static constexpr int E = 78; };
template <> struct RootLog2<6> { static constexpr int E = 39; };

// 2^e mod p at compile time
constexpr u64 pow2_mod_p(int e) {
  u64 r = 1;
  for (int i = 0; i < e; i++) {
    unsigned __int128 d = (unsigned __int128)r * 2;
    r = (u64)(d >= P ? d - P : d);
  }
  return r;
}

// (u - v) * 2^K for a compile-time K in [0, 192): 2^96 = -1 turns K >= 96 into the swapped difference.
// Two formulations (TVM_NTT_SHIFT_TWIDDLES selects the first): shift-and-reduce runs entirely on the ALU pipe
// (~14 instructions); a Montgomery multiplication by the compile-time constant 2^K R costs 17 but only 11 of them on
// the ALU pipe, which is the one that binds the NTT kernels (ncu: pipe_alu 60 %), the 4 IMAD.WIDE go to the FMA pipe.
template <int K>
TVM_HD u64 sub_mul_2k(u64 u, u64 v) {
  static_assert(K >= 0 && K < 192, "K out of range");
  constexpr int J = K >= 96 ? K - 96 : K;
  const u64 d = K >= 96 ? fsub(v, u) : fsub(u, v);
  if constexpr (J == 0) return d;
#ifdef TVM_NTT_SHIFT_TWIDDLES
  else return fmul_2k<J>(d);
#else
  else {
    constexpr u64 C = pow2_mod_p(J + 64);   // Montgomery form of 2^J
    return fmul(d, C);
  }
#endif
}

constexpr int bitrev_c(int x, int bits) {
  int r = 0;
  for (int i = 0; i < bits; i++) r |= ((x >> i) & 1) << (bits - 1 - i);
  return r;
}

// One DIF stage over blocks of 2*HALF elements of e[0..N): (u, v) -> (u + v, (u - v) w_{2 HALF}^{+-j}).
template <int LOGN, int STAGE, bool INV, int I>
struct DifStage {
  template <class Arr>
  static TVM_HD void run(Arr &e) {
    constexpr int N = 1 << LOGN;
    constexpr int HALF = N >> (STAGE + 1);
    constexpr int LOGM = LOGN - STAGE;                      // this stage's butterflies belong to 2^LOGM-point DFTs
    constexpr int blk = I / HALF, j = I % HALF, i0 = blk * 2 * HALF + j, i1 = i0 + HALF;
    constexpr int E = RootLog2<LOGM>::E;
    constexpr int K = INV ? (192 - (E * j) % 192) % 192 : (E * j) % 192;
    const u64 u = e[i0], v = e[i1];
    e[i0] = fadd(u, v);
    e[i1] = sub_mul_2k<K>(u, v);
    DifStage<LOGN, STAGE, INV, I + 1>::run(e);
  }
};
template <int LOGN, int STAGE, bool INV>
struct DifStage<LOGN, STAGE, INV, (1 << LOGN) / 2> {
  template <class Arr>
  static TVM_HD void run(Arr &) {}
};
template <int LOGN, bool INV, int STAGE>
struct DifAll {
  template <class Arr>
  static TVM_HD void run(Arr &e) {
    DifStage<LOGN, STAGE, INV, 0>::run(e);
    DifAll<LOGN, INV, STAGE + 1>::run(e);
  }
};
template <int LOGN, bool INV>
struct DifAll<LOGN, INV, LOGN> {
  template <class Arr>
  static TVM_HD void run(Arr &) {}
};

// In-place 2^LOGN-point DFT of e[0 .. 2^LOGN) with root w_{2^LOGN}^(+-1) (no 1/N scaling).
// Natural order in; OUTPUT IN BIT-REVERSED POSITION: X[k] is left in e[bitrev(k)] — callers index with
// bitrev_c, which costs nothing once the loops are unrolled.
template <int LOGN, bool INV, class Arr>
TVM_HD void dft_pow2(Arr &e) {
  DifAll<LOGN, INV, 0>::run(e);
}

}  // namespace tvm

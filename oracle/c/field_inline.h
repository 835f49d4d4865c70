/* Inline field arithmetic for the generated oracle AIR evaluator.  TEST INFRASTRUCTURE ONLY. */
#ifndef ORC_FIELD_INLINE_H
#define ORC_FIELD_INLINE_H
#include <stdint.h>
typedef uint64_t u64;
typedef unsigned __int128 u128;
#define ORC_P 0xFFFFFFFF00000001ULL
typedef struct { u64 c0, c1, c2; } xfe;
static inline u64 montyred_(u128 x) {
  u64 xl = (u64)x, xh = (u64)(x >> 64);
  u64 a; unsigned e = __builtin_add_overflow(xl, xl << 32, &a);
  u64 b = a - (a >> 32) - e;
  u64 r; unsigned c = __builtin_sub_overflow(xh, b, &r);
  return r - (0xFFFFFFFFULL * c);
}
static inline u64 fmul(u64 a, u64 b) { return montyred_((u128)a * b); }
static inline u64 fadd(u64 a, u64 b) { u64 s; unsigned c = __builtin_add_overflow(a, b, &s); if (c || s >= ORC_P) s -= ORC_P; return s; }
static inline u64 fsub(u64 a, u64 b) { u64 d; unsigned c = __builtin_sub_overflow(a, b, &d); if (c) d += ORC_P; return d; }
static inline xfe xadd(xfe a, xfe b) { xfe r = {fadd(a.c0, b.c0), fadd(a.c1, b.c1), fadd(a.c2, b.c2)}; return r; }
static inline xfe xaddb(xfe a, u64 b) { xfe r = {fadd(a.c0, b), a.c1, a.c2}; return r; }
static inline xfe xmulb(xfe a, u64 b) { xfe r = {fmul(a.c0, b), fmul(a.c1, b), fmul(a.c2, b)}; return r; }
static inline xfe xmul(xfe a, xfe b) {
  u64 d0 = fmul(a.c0, b.c0);
  u64 d1 = fadd(fmul(a.c0, b.c1), fmul(a.c1, b.c0));
  u64 d2 = fadd(fadd(fmul(a.c0, b.c2), fmul(a.c1, b.c1)), fmul(a.c2, b.c0));
  u64 d3 = fadd(fmul(a.c1, b.c2), fmul(a.c2, b.c1));
  u64 d4 = fmul(a.c2, b.c2);
  xfe r = {fsub(d0, d3), fsub(fadd(d1, d3), d4), fadd(d2, d4)};
  return r;
}
#endif

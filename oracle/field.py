"""B-field / X-field arithmetic on Python ints (canonical representation).

Restates `twenty_first::math::{b_field_element, x_field_element}` (crate
`twenty-first = "2.0.0"`, reference `Cargo.toml:104`; not vendored) as used by
the reference at e.g. `triton-vm/src/stark.rs:25`, `arithmetic_domain.rs:7-9`.

* p = 2^64 - 2^32 + 1                         (reference `triton-vm/src/lib.rs:5-6`)
* Montgomery radix R = 2^64, `bfe!(42).raw_u64() == 180388626390`
                                               (`triton-constraint-builder/src/codegen.rs:926-932`)
* X-field = F_p[X]/(X^3 - X + 1)              (`specification/src/isa.md:8`)
* multiplicative generator 7, 2-adic roots = 1753635133440165772^(2^(32-k))
  (twenty-first's PRIMITIVE_ROOTS table; recalled, consistency-checked in
  tests/test_oracle_field.py)

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
P = (1 << 64) - (1 << 32) + 1
R = (1 << 64) % P                 # 2^32 - 1
R_INV = pow(R, P - 2, P)
GENERATOR = 7
ROOT_2_32 = 1753635133440165772   # primitive 2^32-th root of unity


def to_mont(x):   return (x * R) % P
def from_mont(x): return (x * R_INV) % P
def inv(x):
    assert x % P != 0
    return pow(x, P - 2, P)


def primitive_root_of_unity(n):
    """BFieldElement::primitive_root_of_unity(n) for n a power of two <= 2^32."""
    assert n >= 1 and n & (n - 1) == 0 and n <= 1 << 32
    log2n = n.bit_length() - 1
    return pow(ROOT_2_32, 1 << (32 - log2n), P)


def batch_inversion(xs):
    n = len(xs)
    if n == 0: return []
    pre = [1] * n
    acc = 1
    for i, x in enumerate(xs):
        pre[i] = acc
        acc = acc * x % P
    acc = inv(acc)
    out = [0] * n
    for i in range(n - 1, -1, -1):
        out[i] = acc * pre[i] % P
        acc = acc * xs[i] % P
    return out


# ---- X-field: tuples (c0, c1, c2), little-endian coefficients (stark.rs:427-433) ----
X_ZERO = (0, 0, 0)
X_ONE = (1, 0, 0)

def xlift(b): return (b % P, 0, 0)
def xadd(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P, (a[2] + b[2]) % P)
def xsub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P, (a[2] - b[2]) % P)
def xneg(a): return ((-a[0]) % P, (-a[1]) % P, (-a[2]) % P)
def xscale(a, s): return (a[0] * s % P, a[1] * s % P, a[2] * s % P)

def xmul(a, b):
    # X^3 = X - 1  =>  X^4 = X^2 - X
    d0 = a[0] * b[0]
    d1 = a[0] * b[1] + a[1] * b[0]
    d2 = a[0] * b[2] + a[1] * b[1] + a[2] * b[0]
    d3 = a[1] * b[2] + a[2] * b[1]
    d4 = a[2] * b[2]
    return ((d0 - d3) % P, (d1 + d3 - d4) % P, (d2 + d4) % P)

def xpow(a, e):
    r = X_ONE
    while e:
        if e & 1: r = xmul(r, a)
        a = xmul(a, a)
        e >>= 1
    return r

def xinv(a):
    """Inverse in F_p[X]/(X^3-X+1) via the adjugate of the multiplication matrix."""
    assert a != X_ZERO
    # columns of M are a*1, a*X, a*X^2
    c0 = a
    c1 = xmul(a, (0, 1, 0))
    c2 = xmul(a, (0, 0, 1))
    m = [[c0[0], c1[0], c2[0]], [c0[1], c1[1], c2[1]], [c0[2], c1[2], c2[2]]]
    # solve M y = e0 by Cramer
    def det3(q):
        return (q[0][0] * (q[1][1] * q[2][2] - q[1][2] * q[2][1])
                - q[0][1] * (q[1][0] * q[2][2] - q[1][2] * q[2][0])
                + q[0][2] * (q[1][0] * q[2][1] - q[1][1] * q[2][0])) % P
    d = det3(m)
    di = inv(d)
    out = []
    for k in range(3):
        mk = [row[:] for row in m]
        for r in range(3):
            mk[r][k] = 1 if r == 0 else 0
        out.append(det3(mk) * di % P)
    return tuple(out)

def xbatch_inversion(xs):
    n = len(xs)
    if n == 0: return []
    pre = [X_ONE] * n
    acc = X_ONE
    for i, x in enumerate(xs):
        pre[i] = acc
        acc = xmul(acc, x)
    acc = xinv(acc)
    out = [X_ZERO] * n
    for i in range(n - 1, -1, -1):
        out[i] = xmul(acc, pre[i])
        acc = xmul(acc, xs[i])
    return out

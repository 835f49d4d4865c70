"""Naive and fast NTT on Python ints, canonical representation.  Restates the
convention fixed by reference arithmetic_domain.rs:227-246,361-392,457-473:
`domain.evaluate(f)[i] == f(offset * g^i)`, natural order, g = primitive_root_of_unity(len).
TEST INFRASTRUCTURE ONLY."""
from .field import P, primitive_root_of_unity, inv


def naive_evaluate(coef, offset, n):
    g = primitive_root_of_unity(n)
    out = []
    for i in range(n):
        x = offset * pow(g, i, P) % P
        acc = 0
        for c in reversed(coef):
            acc = (acc * x + c) % P
        out.append(acc)
    return out


def ntt(x, omega=None):
    n = len(x)
    if n == 1:
        return list(x)
    if omega is None:
        omega = primitive_root_of_unity(n)
    even = ntt(x[0::2], omega * omega % P)
    odd = ntt(x[1::2], omega * omega % P)
    out = [0] * n
    w = 1
    for k in range(n // 2):
        t = w * odd[k] % P
        out[k] = (even[k] + t) % P
        out[k + n // 2] = (even[k] - t) % P
        w = w * omega % P
    return out


def intt(x):
    n = len(x)
    y = ntt(x, inv(primitive_root_of_unity(n)))
    ni = inv(n)
    return [v * ni % P for v in y]


def coset_evaluate(coef, offset, n):
    """arithmetic_domain.rs:141-170 incl. chunk folding for len(coef) > n."""
    out = [0] * n
    for ch in range(0, max(len(coef), 1), n):
        chunk = list(coef[ch:ch + n])
        scaled = [c * pow(offset, i, P) % P for i, c in enumerate(chunk)] + [0] * (n - len(chunk))
        ev = ntt(scaled)
        so = pow(offset, ch, P)
        out = [(o + e * so) % P for o, e in zip(out, ev)]
    return out


def coset_interpolate(vals, offset):
    c = intt(vals)
    oi = inv(offset)
    return [v * pow(oi, i, P) % P for i, v in enumerate(c)]

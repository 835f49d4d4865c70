"""Python model of the round-2 GPU NTT (csrc/ntt_tile.cu): two passes of square two-round tiles (R x R points per tile row,
R = 2^LOGR points per thread and round, all in-register twiddles powers of two), pre-scale table, geometric post factors.
Validates the index algebra and the exponents against the oracle's definitions; the in-register DFTs are modelled as
plain DFTs (csrc/ntt_radix.cuh has its own host test)."""
import sys, random
sys.path.insert(0, '.')
from oracle.field import P, primitive_root_of_unity as root, inv
from oracle import ntt as N


def dft(x, w):
    n = len(x)
    return [sum(x[j] * pow(w, j * k, P) for j in range(n)) % P for k in range(n)]


def tile_rows(rows, R, wM, pre=None, post=None):
    """rows: list of T lists of M = R*R values.  Models the CTA: thread (t, a) round 1, exchange, thread (t, k2) round 2.
    pre(t, j, v) -> v ; post(t, K, v) -> v.  Returns rows of outputs X[K]."""
    M = R * R
    T = len(rows)
    ex = {}
    for t in range(T):
        for a in range(R):
            v = [rows[t][a + R * b] for b in range(R)]
            if pre:
                v = [pre(t, a + R * b, v[b]) for b in range(R)]
            z = dft(v, pow(wM, R, P))                       # w_R = w_M^R
            for k2 in range(R):
                ex[(k2, a, t)] = z[k2] * pow(wM, a * k2, P) % P
    out = [[0] * M for _ in range(T)]
    for t in range(T):
        for k2 in range(R):
            y = [ex[(k2, a, t)] for a in range(R)]
            x = dft(y, pow(wM, R, P))
            for k1 in range(R):
                K = k2 + R * k1
                out[t][K] = post(t, K, x[k1]) if post else x[k1]
    return out


def evaluate(coef, n, log_r, cosets, RA, RB, T, fold_count):
    """coef: pre-scaled coefficients c'_j, j < n + fold_count.  Returns {c: [n values]} = evaluations on coset c."""
    n2, n1 = RA * RA, RB * RB
    assert n1 * n2 == n
    r = 1 << log_r
    wrn = root(r * n)
    res = {}
    for c in cosets:
        ff = pow(wrn, c * n, P)                              # w_r^c
        S = [pow(wrn, c * n1 * j2, P) for j2 in range(n2)]   # prescale table of this coset
        tmp = [0] * n
        for j1_0 in range(0, n1, T):
            rows = [[coef[j1_0 + t + n1 * j2] for j2 in range(n2)] for t in range(T)]
            def pre(t, j2, v, j1_0=j1_0):
                j = j1_0 + t + n1 * j2
                if j < fold_count:
                    v = (v + ff * coef[n + j]) % P
                return v * S[j2] % P
            def post(t, K, v, j1_0=j1_0):
                j1 = j1_0 + t
                return v * pow(wrn, (j1 * (r * K + c)) % (r * n), P) % P
            Y = tile_rows(rows, RA, root(n2), pre, post)
            for t in range(T):
                for K in range(n2):
                    tmp[K * n1 + j1_0 + t] = Y[t][K]
        out = [0] * n
        for K0 in range(0, n2, T):
            rows = [[tmp[(K0 + t) * n1 + j1] for j1 in range(n1)] for t in range(T)]
            Z = tile_rows(rows, RB, root(n1))
            for t in range(T):
                for KB in range(n1):
                    out[K0 + t + n2 * KB] = Z[t][KB]
        res[c] = out
    return res


def interpolate(vals, n, RA, RB, T, offset):
    """trace column -> pre-scaled coefficients coeff_k * offset^k"""
    n2, n1 = RA * RA, RB * RB
    wi = inv(root(n))
    tmp = [0] * n
    for j1_0 in range(0, n1, T):
        rows = [[vals[j1_0 + t + n1 * j2] for j2 in range(n2)] for t in range(T)]
        def post(t, K, v, j1_0=j1_0):
            return v * pow(wi, ((j1_0 + t) * K) % n, P) % P
        Y = tile_rows(rows, RA, inv(root(n2)), None, post)
        for t in range(T):
            for K in range(n2):
                tmp[K * n1 + j1_0 + t] = Y[t][K]
    out = [0] * n
    ninv = inv(n)
    for K0 in range(0, n2, T):
        rows = [[tmp[(K0 + t) * n1 + j1] for j1 in range(n1)] for t in range(T)]
        def post(t, KB, v, K0=K0):
            k = K0 + t + n2 * KB
            return v * ninv % P * pow(offset, k, P) % P
        Z = tile_rows(rows, RB, inv(root(n1)), None, post)
        for t in range(T):
            for KB in range(n1):
                out[K0 + t + n2 * KB] = Z[t][KB]
    return out


def main():
    rnd = random.Random(1)
    for RA, RB in ((4, 4), (2, 4), (4, 2)):
        n = RA * RA * RB * RB
        T = 2
        offset = 7
        h = 5
        trace = [rnd.randrange(P) for _ in range(n)]
        rand = [rnd.randrange(P) for _ in range(h)]
        # reference: interpolant + zerofier * randomizer, evaluated on offset * <w_{8n}>
        coeff = N.intt(trace)
        full = coeff + [0] * n
        for k in range(h):
            full[k] = (full[k] - rand[k]) % P
            full[n + k] = (full[n + k] + rand[k]) % P
        want_all = N.coset_evaluate(full, offset, 8 * n)
        got_coef = interpolate(trace, n, RA, RB, T, offset)
        assert got_coef == [c * pow(offset, k, P) % P for k, c in enumerate(coeff)], "interpolate"
        pre_scaled = [c * pow(offset, k, P) % P for k, c in enumerate(full)][:n + h]
        ev = evaluate(pre_scaled, n, 3, range(8), RA, RB, T, h)
        for c in range(8):
            assert ev[c] == [want_all[c + 8 * k] for k in range(n)], ("evaluate", c)
        print("ok", RA, RB)


if __name__ == "__main__":
    main()

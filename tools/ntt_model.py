"""Python model of the GPU NTT decomposition (4-step, radix-8 DIF tiles, bit-reversed row
addressing, inter-pass twist) — used to validate the index algebra before writing CUDA."""
import sys, random
sys.path.insert(0, '.')
from oracle.field import P, primitive_root_of_unity, inv
from oracle import ntt as N

def bitrev(x, bits):
    r = 0
    for i in range(bits):
        r |= ((x >> i) & 1) << (bits - 1 - i)
    return r

def dft8net(e, W8):
    W = [pow(W8, i, P) for i in range(4)]
    for i in range(4):
        u, v = e[i], e[i + 4]
        e[i] = (u + v) % P; e[i + 4] = (u - v) * W[i] % P
    for h in (0, 4):
        for i in range(2):
            u, v = e[h + i], e[h + i + 2]
            e[h + i] = (u + v) % P; e[h + i + 2] = (u - v) * W[2 * i] % P
    for q in (0, 2, 4, 6):
        u, v = e[q], e[q + 1]
        e[q] = (u + v) % P; e[q + 1] = (u - v) % P

def tile_dif(row, LT, wN):
    """in-place DIF of one row of size 2^LT, natural in -> bit-reversed out"""
    Nn = 1 << LT
    tw = [pow(wN, e, P) for e in range(Nn)]
    W8 = pow(wN, Nn // 8, P) if LT >= 3 else None
    W4 = pow(wN, Nn // 4, P) if LT >= 2 else None
    m_log = LT
    while m_log >= 3:
        for g in range(Nn // 8):
            j0 = g & ((1 << (m_log - 3)) - 1)
            b = (g >> (m_log - 3)) << m_log
            pos = [b + j0 + (i << (m_log - 3)) for i in range(8)]
            e = [row[p] for p in pos]
            dft8net(e, W8)
            for i in range(8):
                ex = (j0 * bitrev(i, 3)) << (LT - m_log)
                row[pos[i]] = e[i] * tw[ex] % P
        m_log -= 3
    if m_log == 2:
        for g in range(Nn // 4):
            p = 4 * g
            e = row[p:p + 4]
            for i in range(2):
                u, v = e[i], e[i + 2]
                e[i] = (u + v) % P; e[i + 2] = (u - v) * pow(W4, i, P) % P
            for q in (0, 2):
                u, v = e[q], e[q + 1]
                e[q] = (u + v) % P; e[q + 1] = (u - v) % P
            row[p:p + 4] = e
    elif m_log == 1:
        for g in range(Nn // 2):
            u, v = row[2 * g], row[2 * g + 1]
            row[2 * g] = (u + v) % P; row[2 * g + 1] = (u - v) % P

def ntt_two_pass(x, L, LA, inverse=False):
    """n = 2^L = n1*n2, n2 = 2^LA (pass A tile), n1 = 2^(L-LA) (pass B tile)."""
    n = 1 << L; LB = L - LA; n1 = 1 << LB; n2 = 1 << LA
    wn = primitive_root_of_unity(n)
    if inverse: wn = inv(wn)
    wn2 = pow(wn, n1, P); wn1 = pow(wn, n2, P)
    Y = [0] * n
    for j1 in range(n1):
        row = [x[j1 + n1 * j2] for j2 in range(n2)]
        tile_dif(row, LA, wn2)
        for pos in range(n2):
            k2 = bitrev(pos, LA)
            Y[j1 + n1 * k2] = row[pos] * pow(wn, j1 * k2, P) % P
    X = [0] * n
    for k2 in range(n2):
        row = [Y[j1 + n1 * k2] for j1 in range(n1)]
        tile_dif(row, LB, wn1)
        for pos in range(n1):
            k1 = bitrev(pos, LB)
            X[k2 + n2 * k1] = row[pos]
    if inverse:
        ni = inv(n)
        X = [v * ni % P for v in X]
    return X

if __name__ == "__main__":
    random.seed(3)
    for L, LA in [(6, 3), (7, 4), (8, 4), (9, 5), (10, 4), (5, 0), (4, 2), (3, 1), (11, 6)]:
        x = [random.randrange(P) for _ in range(1 << L)]
        assert ntt_two_pass(x, L, LA) == N.ntt(x), (L, LA)
        assert ntt_two_pass(x, L, LA, True) == N.intt(x), (L, LA)
    print("model ok")
    # the specific constants used on device: w8 = -2^24 etc.
    w8 = primitive_root_of_unity(8); w16 = primitive_root_of_unity(16)
    assert w8 == P - (1 << 24) and pow(w8, 2, P) == 1 << 48 and pow(w8, 3, P) == P - pow(2, 72, P)
    print("w8 ok", w16 == P - (1 << 60))

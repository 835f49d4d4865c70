"""GPU parity tests (through the C ABI) for the NTT / LDE / Tip5 / Merkle kernels against the
CPU oracle on the same seeded inputs.  Bit-exact: integer arithmetic."""
import numpy as np
import pytest

from conftest import rand_bfes
from oracle import corc, field as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("log2n", [0, 1, 2, 3, 4, 5, 6, 7, 9, 10, 12, 13, 14, 16, 17, 20, 25])
@pytest.mark.parametrize("inverse", [False, True])
def test_ntt_matches_oracle(backend, log2n, inverse):
    rng = np.random.default_rng(100 + log2n)
    ncols = 1 if log2n >= 20 else 3 if log2n >= 13 else 5
    x = rand_bfes(rng, (ncols, 1 << log2n))
    got = backend.ntt(x, inverse=inverse)
    for c in range(ncols):
        assert np.array_equal(got[c], corc.ntt(x[c], inverse=inverse)), (log2n, inverse, c)


def test_ntt_edge_values(backend):
    # all-zero, all p-1, delta
    n = 1 << 10
    x = np.zeros((3, n), dtype=np.uint64)
    x[1, :] = F.P - 1
    x[2, 0] = 1
    got = backend.ntt(x)
    assert not got[0].any()
    assert np.array_equal(got[1], corc.ntt(x[1]))
    assert (got[2] == 1).all()


def test_ntt_roundtrip_large(backend):
    # size-independent property at BASELINE scale: intt(ntt(x)) == x for 2^20
    rng = np.random.default_rng(5)
    x = rand_bfes(rng, (2, 1 << 20))
    y = backend.ntt(x)
    assert np.array_equal(y[0], corc.ntt(x[0]))
    assert np.array_equal(backend.ntt(y, inverse=True), x)


@pytest.mark.parametrize("log2_trace,log2_cosets,h,ncols", [(4, 3, 5, 3), (8, 3, 40, 4), (10, 3, 198, 7), (13, 3, 240, 3),
                                                             (6, 2, 0, 2), (12, 1, 100, 2), (3, 3, 8, 2)])
def test_lde_matches_oracle(backend, log2_trace, log2_cosets, h, ncols):
    rng = np.random.default_rng(7 * log2_trace + h)
    n = 1 << log2_trace
    trace = rand_bfes(rng, (ncols, n))
    rand = rand_bfes(rng, (ncols, h)) if h else None
    got = backend.lde(trace, rand, log2_cosets, 7)
    want = corc.lde_table(trace, rand, 7, log2_trace + log2_cosets)
    assert np.array_equal(got, want)


def test_lde_subsampling_property(backend):
    # arithmetic_domain.rs:395-415 + zero-knowledge.md: randomizers vanish on the trace domain.
    # Evaluating with offset 1 on r cosets: rows i = r*k are the trace itself.
    rng = np.random.default_rng(3)
    n, h = 1 << 14, 200
    trace = rand_bfes(rng, (2, n))
    rand = rand_bfes(rng, (2, h))
    got = backend.lde(trace, rand, 2, 1)
    assert np.array_equal(got[:, ::4], trace)


@pytest.mark.parametrize("ncols", [1, 9, 10, 11, 15, 20, 91, 379])
def test_hash_rows_matches_oracle(backend, ncols):
    rng = np.random.default_rng(ncols)
    nrows = 300  # ragged vs the 128-thread CTA
    tab = rand_bfes(rng, (ncols, nrows))
    got = backend.hash_rows(tab)
    assert np.array_equal(got, corc.hash_rows_colmajor(tab))


def test_hash_rows_extreme_values(backend):
    tab = np.zeros((12, 64), dtype=np.uint64)
    tab[:, 1::2] = F.P - 1
    got = backend.hash_rows(tab)
    assert np.array_equal(got, corc.hash_rows_colmajor(tab))


@pytest.mark.parametrize("log2_leaves", [0, 1, 2, 5, 7, 8, 12, 15])
def test_merkle_matches_oracle(backend, log2_leaves):
    rng = np.random.default_rng(log2_leaves)
    leaves = rand_bfes(rng, (1 << log2_leaves, 5))
    root, nodes = backend.merkle(leaves, want_nodes=True)
    want = corc.merkle_build(leaves)
    assert np.array_equal(nodes[1:], want[1:])
    assert np.array_equal(root, want[1])


def test_device_pipeline_lde_hash_merkle(backend):
    """HBM-resident path: trace -> LDE (coset-major) -> row digests (natural order) -> root."""
    import torch
    rng = np.random.default_rng(11)
    log2_trace, log2_cosets, ncols, h = 10, 3, 23, 50
    n, rn = 1 << log2_trace, 1 << (log2_trace + log2_cosets)
    trace = rand_bfes(rng, (ncols, n))
    rand = rand_bfes(rng, (ncols, h))
    dev = torch.device("cuda:0")
    d_trace = torch.from_numpy(trace.view(np.int64)).to(dev)
    d_rand = torch.from_numpy(rand.view(np.int64)).to(dev)
    backend.to_mont_(d_trace); backend.to_mont_(d_rand)
    d_coef = torch.empty((ncols, 2 * n), dtype=torch.int64, device=dev)
    d_out = torch.empty((ncols, rn), dtype=torch.int64, device=dev)
    d_tmp = torch.empty((ncols, rn), dtype=torch.int64, device=dev)
    backend.lde_dev(d_trace, d_rand, h, log2_trace, log2_cosets, 7, ncols, d_coef, d_out, d_tmp)
    d_nodes = torch.zeros((2 * rn, 5), dtype=torch.int64, device=dev)
    backend.hash_rows_dev(d_out, rn, rn, ncols, log2_cosets, d_nodes[rn:])
    backend.merkle_dev(d_nodes, rn)
    backend.from_mont_(d_nodes)
    backend.synchronize()
    nodes = d_nodes.cpu().numpy().view(np.uint64)
    lde = corc.lde_table(trace, rand, 7, log2_trace + log2_cosets)
    digests = corc.hash_rows_colmajor(lde)
    want = corc.merkle_build(digests)
    assert np.array_equal(nodes[rn:], digests)
    assert np.array_equal(nodes[1], want[1])
    assert backend.launches > 0


def test_air_quotient_matches_python_evaluator(backend):
    """Generated quotient kernels vs node-by-node evaluation of the same circuits in Python
    (which is pinned to the reference's golden fingerprint in tests/test_air.py)."""
    import os
    import sys
    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "triton-vm_b200"))
    from airgen.build import CATEGORIES, build_air
    from airgen.evaluate import evaluate_constraints
    air = build_air()
    rng = np.random.default_rng(77)
    log_n, log_r = 4, 3
    n, rn = 1 << log_n, 1 << (log_n + log_r)
    r = 1 << log_r
    main = rand_bfes(rng, (379, rn))          # memory order: [col][coset*n + k]
    aux = rand_bfes(rng, (270, rn))
    ch = rand_bfes(rng, (63, 3))
    w = rand_bfes(rng, (604, 3))
    dev = torch.device("cuda:0")
    d_main = torch.from_numpy(main.view(np.int64)).to(dev)
    d_aux = torch.from_numpy(aux.view(np.int64)).to(dev)
    backend.to_mont_(d_main); backend.to_mont_(d_aux)
    d_out = torch.zeros((3, rn), dtype=torch.int64, device=dev)
    backend.air_quotient_dev(d_main, rn, d_aux, rn, ch, w, log_n, log_r, 7, d_out, rn)
    backend.from_mont_(d_out)
    backend.synchronize()
    got = d_out.cpu().numpy().view(np.uint64)

    P = F.P
    wn = F.primitive_root_of_unity(n)
    wrn = F.primitive_root_of_unity(rn)
    wn_inv = F.inv(wn)
    chal = [tuple(int(v) for v in row) for row in ch]
    weights = [tuple(int(v) for v in row) for row in w]
    for m in list(range(0, rn, 7)) + [n - 1, rn - 1]:
        c, k = m // n, m % n
        m_next = c * n + (k + 1) % n
        x = 7 * pow(wrn, c + r * k, P) % P
        cm = [int(v) for v in main[:, m]]; nm = [int(v) for v in main[:, m_next]]
        ca = [tuple(int(v) for v in aux[3 * q:3 * q + 3, m]) for q in range(90)]
        na = [tuple(int(v) for v in aux[3 * q:3 * q + 3, m_next]) for q in range(90)]
        zinv = {"init": F.inv((x - 1) % P), "cons": F.inv((pow(x, n, P) - 1) % P),
                "tran": (x - wn_inv) * F.inv((pow(x, n, P) - 1) % P) % P, "term": F.inv((x - wn_inv) % P)}
        acc, off = (0, 0, 0), 0
        for cat in CATEGORIES:
            vals = evaluate_constraints(air.constraints[cat], cm, ca, nm, na, chal)
            s = (0, 0, 0)
            for j, v in enumerate(vals):
                s = F.xadd(s, F.xmul(weights[off + j], v))
            off += len(vals)
            acc = F.xadd(acc, F.xscale(s, zinv[cat]))
        assert tuple(int(got[d, m]) for d in range(3)) == acc, m

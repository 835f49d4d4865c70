"""GPU parity tests (through the C ABI) for the NTT / LDE / Tip5 / Merkle kernels against the
CPU oracle on the same seeded inputs.  Bit-exact: integer arithmetic."""
import numpy as np
import pytest

from conftest import rand_bfes
from oracle import corc, field as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("log2n", [0, 1, 2, 3, 4, 5, 6, 7, 9, 10, 12, 13, 14, 16, 17])
@pytest.mark.parametrize("inverse", [False, True])
def test_ntt_matches_oracle(backend, log2n, inverse):
    rng = np.random.default_rng(100 + log2n)
    ncols = 3 if log2n >= 13 else 5
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

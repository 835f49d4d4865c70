"""The reference's STIR property tests (low_degree_test/stir.rs:1813-2010) restated against this repository's STIR: the native
verifier `tvm_stir_verify` (host code, runs everywhere) and, on a GPU, the device prover `tvm_stir_prove` — for ARBITRARY
StirParameters, not only the ones Stark::default() derives.  The oracle (oracle/stir.py) is the checker: its prover feeds the
CPU tests, its proof words are what the device prover must reproduce."""
import random

import numpy as np
import pytest

from oracle import codec, field as F, stir as ST
from oracle.field import P

PARAMS = [(8, 1, 5), (12, 2, 6), (20, 2, 8), (16, 3, 7), (40, 2, 9), (10, 1, 9), (30, 4, 6)]   # (security, log2 expansion, log2 high-degree bound)


def _instance(security, log2_exp, log2_hdb, degree, seed):
    rnd = random.Random(seed)
    sp = ST.derive(security, 2, log2_exp, log2_hdb)
    poly = [tuple(rnd.randrange(P) for _ in range(3)) for _ in range(degree + 1)]
    cw = ST.xevaluate(poly, sp["initial_offset"], sp["initial_domain_len"]) if poly else np.zeros((sp["initial_domain_len"], 3), dtype=np.uint64)
    return sp, [tuple(int(t) for t in v) for v in cw]


def _oracle_proof(sp, cw):
    ps = codec.ProofStream()
    idx = ST.prove(ps, cw, sp)
    return ps.encode(), idx


@pytest.mark.parametrize("security,log2_exp,log2_hdb", PARAMS)
def test_native_verifier_accepts_low_degree_for_arbitrary_parameters(security, log2_exp, log2_hdb):
    """prove_and_verify_low_degree_polynomial (stir.rs:1869-1900): accepted, same first-round indices, the partial first
    codeword is the codeword at those indices; prove_and_verify_zero_polynomial (1813-1822)"""
    import tvm_b200
    max_degree = (1 << log2_hdb) - 1
    for degree in (max_degree, max_degree // 3, 0, -1):
        sp, cw = _instance(security, log2_exp, log2_hdb, degree, seed=degree + 7)
        proof, idx = _oracle_proof(sp, cw)
        ok, why, vidx, vals = tvm_b200.stir_verify(security, log2_exp, log2_hdb, proof)
        assert ok, why
        assert vidx == idx
        assert [tuple(int(t) for t in v) for v in vals] == [cw[i] for i in idx]


@pytest.mark.parametrize("security,log2_exp,log2_hdb", PARAMS[:5])
def test_native_verifier_rejects_high_degree(security, log2_exp, log2_hdb):
    """prove_and_fail_to_verify_high_degree_polynomial (stir.rs:1903-1920)"""
    import tvm_b200
    too_high = 1 << log2_hdb
    for degree in (too_high, too_high + too_high // 2, 2 * too_high - 1):
        sp, cw = _instance(security, log2_exp, log2_hdb, degree, seed=degree)
        proof, _ = _oracle_proof(sp, cw)
        ok, why, _, _ = tvm_b200.stir_verify(security, log2_exp, log2_hdb, proof)
        assert not ok and why


def _item_spans(proof):
    """(kind, first payload word, end) of every item of an encoded proof stream"""
    spans, pos = [], 2
    for _ in range(int(proof[1])):
        ln = int(proof[pos]); start = pos + 1
        spans.append((int(proof[start]), start + 1, start + ln))
        pos = start + ln
    return spans


@pytest.mark.parametrize("security,log2_exp,log2_hdb", [(12, 2, 6), (20, 2, 8), (8, 2, 10)])
def test_modified_proof_stream_results_in_verification_failure(security, log2_exp, log2_hdb):
    """modified_proof_stream_results_in_verification_failure (stir.rs:1967-2010): one corrupted word in any item - Merkle
    root, out-of-domain values, queried leafs, authentication structure, final polynomial - is rejected"""
    import tvm_b200
    sp, cw = _instance(security, log2_exp, log2_hdb, (1 << log2_hdb) - 1, seed=5)
    proof, _ = _oracle_proof(sp, cw)
    proof = np.array(proof, dtype=np.uint64)
    assert tvm_b200.stir_verify(security, log2_exp, log2_hdb, proof)[0]
    rnd = random.Random(9)
    kinds = set()
    for kind, a, b in _item_spans(proof):
        for _ in range(3):
            bad = proof.copy()
            j = rnd.randrange(a, b)
            bad[j] = (int(bad[j]) + 1 + rnd.randrange(P - 2)) % P
            ok, why, _, _ = tvm_b200.stir_verify(security, log2_exp, log2_hdb, bad)
            assert not ok, (kind, j - a)
        kinds.add(kind)
    assert len(kinds) >= (4 if sp["round_queries"] else 3)   # MerkleRoot, StirResponse, Polynomial (+ StirOutOfDomainValues with full rounds)
    # truncated / extended streams
    assert not tvm_b200.stir_verify(security, log2_exp, log2_hdb, proof[:-3])[0]
    # parameters the proof was not made for (security level / expansion): rejected, no crash
    assert not tvm_b200.stir_verify(security + 9, log2_exp, log2_hdb, proof)[0]


def test_verifying_arbitrary_words_does_not_crash():
    """verifying_arbitrary_proof_does_not_panic (stir.rs:1960-1963)"""
    import tvm_b200
    rnd = random.Random(1)
    for n in (1, 2, 7, 100, 5000):
        junk = np.array([rnd.randrange(P) for _ in range(n)], dtype=np.uint64)
        assert tvm_b200.stir_verify(12, 2, 6, junk)[0] is False


@pytest.mark.gpu
@pytest.mark.parametrize("security,log2_exp,log2_hdb", PARAMS + [(24, 2, 11), (160, 2, 12)])
def test_device_stir_prover_for_arbitrary_parameters(backend, security, log2_exp, log2_hdb):
    """the device prover under arbitrary StirParameters: word for word the oracle's proof stream, accepted by the native
    verifier with the prover's indices (low degree), rejected for a too-high degree"""
    import tvm_b200
    max_degree = (1 << log2_hdb) - 1
    for degree in (max_degree, max_degree // 2, -1):
        sp, cw = _instance(security, log2_exp, log2_hdb, degree, seed=degree + 3)
        want, want_idx = _oracle_proof(sp, cw)
        got, idx = backend.stir_prove(security, log2_exp, log2_hdb, np.array(cw, dtype=np.uint64))
        assert [int(v) for v in got] == want and idx == want_idx
        ok, why, vidx, vals = tvm_b200.stir_verify(security, log2_exp, log2_hdb, got)
        assert ok and vidx == idx, why
    sp, cw = _instance(security, log2_exp, log2_hdb, max_degree + 1 + max_degree // 2, seed=2)
    got, _ = backend.stir_prove(security, log2_exp, log2_hdb, np.array(cw, dtype=np.uint64))
    assert not tvm_b200.stir_verify(security, log2_exp, log2_hdb, got)[0]

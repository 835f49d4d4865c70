"""CPU tests of the oracle's Stark::prove / verify restatement (oracle/stark.py, codec.py) and of
the C++ parameter derivation exported through the C ABI."""
import numpy as np
import pytest

import tvm_b200
from conftest import rand_bfes
from oracle import codec, field as F, stark as S


def test_parameter_derivation_matches_reference_facts():
    # SURVEY.md §8 table (derived from stark.rs:1885-2089, fri.rs:816-924): default Stark at 2^20 with FRI
    d = S.Stark(160, 2, "fri").derive(1 << 20)
    assert d["num_collinearity_checks"] == 173          # ceil(160 / -log2(0.525))
    assert d["num_trace_randomizers"] == 198            # 173 + 4*3*2 + 1
    assert (d["trace_len"], d["randomized_trace_len"], d["quotient_len"], d["ldt_len"]) == (1 << 20, 1 << 21, 1 << 23, 1 << 23)
    assert d["fri_num_rounds"] == 13 and d["fri_last_round_max_degree"] == 255
    assert d["num_quotient_randomizer_coefficients"] == (198 + 1) * 5
    # Fibonacci(100)-sized instance: padded height 2^10 -> LDT/quotient domain 2^13
    d = S.Stark(160, 2).derive(1 << 10)
    assert (d["trace_len"], d["ldt_len"], d["quotient_len"]) == (1 << 10, 1 << 13, 1 << 13)
    # tiny padded heights are bumped by the number of trace randomizers (stark.rs:1885-1890)
    d = S.Stark(160, 2).derive(4)
    assert d["trace_len"] == 512 and d["ldt_len"] == 4096


@pytest.mark.parametrize("sec,le,ph", [(160, 2, 1 << 20), (160, 2, 1 << 10), (160, 2, 1 << 16), (4, 2, 16), (32, 2, 256),
                                        (160, 2, 1 << 22), (160, 1, 1 << 12), (80, 3, 1 << 9)])
def test_c_abi_derive_domains_matches_oracle(sec, le, ph):
    a = tvm_b200.derive_domains(sec, le, ph, tvm_b200.LDT_FRI)
    b = S.Stark(sec, le, "fri").derive(ph)
    assert all(a[k] == b[k] for k in a if k in b and not k.startswith("stir") and k != "ldt"), (a, b)


def test_invalid_ldt_choices_are_reported():
    with pytest.raises(tvm_b200.TvmError):
        tvm_b200.derive_domains(160, 2, 1 << 16, ldt_choice=3)
    # 0 = the reference's heuristic: STIR from 2^16 on, FRI below
    assert tvm_b200.derive_domains(160, 2, 1 << 16, ldt_choice=0)["ldt"] == tvm_b200.LDT_STIR
    assert tvm_b200.derive_domains(160, 2, 1 << 15, ldt_choice=0)["ldt"] == tvm_b200.LDT_FRI


def test_codec_roundtrip_and_fiat_shamir_flags():
    ps = codec.ProofStream()
    ps.enqueue("Log2PaddedHeight", 10)
    ps.enqueue("MerkleRoot", [1, 2, 3, 4, 5])
    ps.enqueue("FriCodeword", [(1, 2, 3), (4, 5, 6)])
    ps.enqueue("Polynomial", [(1, 0, 0), (0, 0, 7), (0, 0, 0)])      # trailing zero is stripped
    ps.enqueue("FriResponse", ([(9, 8, 7)], [[1, 1, 1, 1, 1], [2, 2, 2, 2, 2]]))
    ps.enqueue("AuthenticationStructure", [[5, 4, 3, 2, 1]])
    words = ps.encode()
    assert words[0] == len(words) - 1 and words[1] == 6
    back = codec.decode_proof(words)
    assert [k for k, _ in back.items] == [k for k, _ in ps.items]
    assert back.items[3][1] == [(1, 0, 0), (0, 0, 7)]
    assert back.items[4][1] == ([(9, 8, 7)], [[1, 1, 1, 1, 1], [2, 2, 2, 2, 2]])
    # items excluded from Fiat-Shamir do not move the sponge (proof_item.rs:96-147)
    a, b = codec.ProofStream(), codec.ProofStream()
    a.enqueue("MerkleRoot", [1, 2, 3, 4, 5]); b.enqueue("MerkleRoot", [1, 2, 3, 4, 5])
    b.enqueue("AuthenticationStructure", [[5, 4, 3, 2, 1]])
    b.enqueue("FriCodeword", [(1, 2, 3)])
    assert a.sponge.state == b.sponge.state
    b.enqueue("Polynomial", [(1, 2, 3)])
    assert a.sponge.state != b.sponge.state


def _instance(seed=1, sec=4, ph=16):
    st = S.Stark(sec, 2)
    d = st.derive(ph)
    rng = np.random.default_rng(seed)
    n, h = d["trace_len"], d["num_trace_randomizers"]
    main, mrand = rand_bfes(rng, (379, n)), rand_bfes(rng, (379, h))
    qrand = rand_bfes(rng, (d["num_quotient_randomizer_coefficients"], 3))

    def aux(_ch):
        r = np.random.default_rng(seed + 100)
        return rand_bfes(r, (91, n, 3)), rand_bfes(r, (91, h, 3))
    return st, d, S.Claim([1, 2, 3, 4, 5], [7], [8, 9]), main, mrand, aux, qrand


def test_oracle_prove_then_verify_structure_and_tamper_detection():
    st, d, claim, main, mrand, aux, qrand = _instance()
    proof, art = S.prove(st, claim, main, mrand, aux, qrand, padded_height=16, keep=True)
    # synthetic traces do not satisfy the AIR: every check except the out-of-domain AIR identity must pass
    assert S.verify(st, claim, proof, check_air=False)
    with pytest.raises(ValueError, match="OutOfDomainQuotientValueMismatch"):
        S.verify(st, claim, proof, check_air=True)
    # a different claim changes the Fiat-Shamir transcript
    with pytest.raises(ValueError):
        S.verify(st, S.Claim([1, 2, 3, 4, 6], [7], [8, 9]), proof, check_air=False)
    bad = list(proof)
    bad[len(bad) // 2] = (bad[len(bad) // 2] + 1) % F.P
    with pytest.raises(ValueError):
        S.verify(st, claim, bad, check_air=False)
    # split-segments recomposition (stark.rs:4432-4501): q(x) = sum_i x^i s_i(x^4) - randomizer part cancels
    # segment codewords are committed as rows of 5 X-field elements
    assert len(art["segment_codewords"]) == 5 and art["segment_codewords"][0].shape == (d["ldt_len"], 3)


def test_quotient_segments_recompose_to_quotient():
    """stark.rs:1252-1263 / zero-knowledge.md: sum_i x^i q_i(x^4) = q(x); the randomized segments satisfy
    sum_i x^i s_i(x^4) + sum_i (zeta x)^i ... at the level the verifier uses (stark.rs:1525-1540)."""
    st, d, claim, main, mrand, aux, qrand = _instance(seed=3)
    _, art = S.prove(st, claim, main, mrand, aux, qrand, padded_height=16, keep=True)
    polys = art["segment_polys"]
    x = (123456789, 987654321, 55555)
    x4 = F.xpow(x, 4)
    zx = F.xscale(x, S.ZETA)
    zx4 = F.xpow(zx, 4)
    lhs = F.X_ZERO
    for i in range(4):
        lhs = F.xadd(lhs, F.xmul(F.xpow(x, i), S.xpoly_eval(polys[i], x4)))
    for i in range(4):
        lhs = F.xadd(lhs, F.xmul(F.xpow(zx, i), S.xpoly_eval(polys[i + 1], zx4)))
    # equals the interpolated quotient polynomial evaluated at x
    qpoly = S.xcoset_interpolate(art["quotient_codeword"], d["ldt_offset"])
    assert lhs == S.xpoly_eval(qpoly, x)

"""STIR: parameter maths pinned against the figures stated in the reference, host C++ derivation == oracle,
oracle prover accepted / corrupted proofs rejected by the restated verifier (CPU only)."""
import numpy as np
import pytest

import tvm_b200
from oracle import codec, field as F, stark as S, stir


def test_log2_binomial_coefficient_table():              # stir.rs:1542-1580
    for want, (a, b) in [(0.0, (10, 0)), (3.322, (10, 1)), (5.492, (10, 2)), (7.977, (10, 5)), (3.322, (10, 9)), (0.0, (10, 10)),
                         (230.424, (500, 50)), (356.476, (500, 100)), (495.191, (500, 250)), (230.424, (500, 450)),
                         (4446.650, (1 << 13, 1 << 10)), (6639.372, (1 << 13, 2 << 10)), (8185.174, (1 << 13, 4 << 10)),
                         (4446.650, (1 << 13, 7 << 10))]:
        assert abs(stir.log2_binomial_coefficient(a, b) - want) < 1e-3


def test_error_margin_examples_from_the_reference_comment():   # stir.rs:738-746
    assert stir.num_total_in_domain_queries(160, 23, 160) == 184
    assert stir.num_total_in_domain_queries(160, 8, 160) == 610


def test_stacking_and_folding_vectors():                 # stir.rs:1528-1539, 1634-1644
    assert stir.stack(list(range(8)), 4) == [[0, 2, 4, 6], [1, 3, 5, 7]]
    assert stir.stack(list(range(8)), 2) == [[0, 4], [1, 5], [2, 6], [3, 7]]
    c = [F.xlift(v) for v in range(1, 11)]
    assert stir.fold_polynomial(c, 4, F.xlift(10)) == [F.xlift(4321), F.xlift(8765), F.xlift(109)]


def test_round_structure_at_the_baseline_heights():      # SURVEY.md §8 table (derived from stir.rs:437-567)
    d = S.Stark(160, 2).derive(1 << 20)
    assert d["ldt"] == "stir" and d["num_trace_randomizers"] == 228 and d["ldt_len"] == 1 << 23
    assert d["stir"]["round_queries"] == [(203, 1), (136, 1), (105, 1), (87, 1), (76, 1), (68, 1)]
    assert d["stir"]["final_num_in_domain_queries"] == 63 and d["stir"]["final_degree"] == 127
    d = S.Stark(160, 2).derive(1 << 16)
    assert d["num_first_round_queries"] == 215 and d["num_trace_randomizers"] == 240
    assert S.Stark(160, 2).derive(1 << 10)["ldt"] == "fri"


def test_q_ary_entropy_reference_values():               # mod.rs:406-423 (computed with sage there)
    for want, e in [(0.505208333333361, 1), (0.254225406898247, 2), (0.127831064808346, 3), (0.064256719096972, 4),
                    (0.032294907939134, 5), (0.016229766017215, 6), (0.008155804230956, 7), (0.004098304720073, 8)]:
        assert abs(stir.rs_q_ary_entropy(e) - want) < 1e-4


@pytest.mark.parametrize("security,log2_exp", [(160, 2), (80, 2), (42, 3), (16, 1), (8, 2), (6, 2)])
@pytest.mark.parametrize("choice", [0, 1, 2])
@pytest.mark.parametrize("soundness", ["proven", "conjectured"])
def test_host_derivation_matches_oracle(security, log2_exp, choice, soundness):
    for log2_ph in (4, 8, 10, 13, 16, 17, 18, 20, 22):
        st = S.Stark(security, log2_exp, {0: None, 1: "fri", 2: "stir"}[choice], soundness)
        want = st.derive(1 << log2_ph)
        got = tvm_b200.derive_domains(security, log2_exp, 1 << log2_ph, choice, conjectured=soundness == "conjectured")
        for k in ("padded_height", "num_trace_randomizers", "randomized_trace_len", "trace_len", "quotient_len", "ldt_len",
                  "num_collinearity_checks", "num_quotient_randomizer_coefficients", "num_first_round_queries"):
            assert got[k] == want[k], (k, log2_ph, got[k], want[k])
        assert got["ldt"] == {"fri": 1, "stir": 2}[want["ldt"]]
        if want["ldt"] == "stir":
            assert got["stir_round_queries"] == want["stir"]["round_queries"]
            assert got["stir_final_num_queries"] == want["stir"]["final_num_in_domain_queries"]
            assert got["stir_final_degree"] == want["stir"]["final_degree"]
        else:
            assert got["fri_num_rounds"] == want["fri_num_rounds"]


@pytest.mark.parametrize("security,hdb", [(42, 11), (10, 13), (8, 6)])
def test_oracle_stir_prove_then_verify(security, hdb):
    sp = stir.derive(security, 2, 2, hdb)
    rng = np.random.default_rng(hdb)
    coeffs = [tuple(int(v) for v in rng.integers(0, F.P, 3, dtype=np.uint64)) for _ in range(1 << hdb)]
    cw = [tuple(int(t) for t in v) for v in stir.xevaluate(coeffs, sp["initial_offset"], sp["initial_domain_len"])]
    ps = codec.ProofStream()
    idx = stir.prove(ps, cw, sp)
    words = ps.encode()
    got_idx, partial = stir.verify(codec.decode_proof(words), sp)
    assert got_idx == idx and partial == [cw[i] for i in idx]
    # a polynomial above the degree bound is rejected
    cw_bad = [tuple(int(t) for t in v) for v in stir.xevaluate(coeffs + [(1, 0, 0)] * 3, sp["initial_offset"], sp["initial_domain_len"])]
    ps = codec.ProofStream()
    stir.prove(ps, cw_bad, sp)
    with pytest.raises(ValueError):
        stir.verify(codec.decode_proof(ps.encode()), sp)
    # so is a corrupted transcript
    bad = list(words)
    bad[len(bad) // 3] = (bad[len(bad) // 3] + 1) % F.P
    with pytest.raises((ValueError, KeyError, IndexError)):
        stir.verify(codec.decode_proof(bad), sp)


def test_ldt_parameter_table_of_the_reference_range():
    """print_various_ldt_parameters (stark.rs:4901-4958): Stark::default() at padded heights 2^8..2^29, both low-degree tests,
    both proximity regimes.  The committed table comes from the oracle (tests/golden/make_ldt_parameter_table.py); the
    product's derivation (tvm_derive_domains) must reproduce every row."""
    import json, os
    table = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ldt_parameter_table.json")))
    assert len(table) == 2 * 2 * 22
    for key, want in table.items():
        ldt, soundness, h = key.split("/")
        got = tvm_b200.derive_domains(160, 2, 1 << int(h), {"fri": 1, "stir": 2}[ldt], conjectured=soundness == "conjectured")
        assert got["ldt_len"].bit_length() - 1 == want["log2_initial_domain_len"], key
        assert got["num_trace_randomizers"] == want["num_trace_randomizers"] and got["trace_len"] == want["trace_len"], key
        if ldt == "fri":
            assert got["fri_num_rounds"] == want["num_rounds"] and got["num_collinearity_checks"] == want["first_round_queries"], key
            assert (got["fri_last_round_max_degree"] + 1).bit_length() - 1 == want["log2_final_degree_plus_1"], key
        else:
            rq = got["stir_round_queries"]
            assert len(rq) == want["num_rounds"] and got["num_first_round_queries"] == want["first_round_queries"], key
            assert sum(a + b for a, b in rq) + got["stir_final_num_queries"] == want["total_queries"], key
            assert (got["stir_final_degree"] + 1).bit_length() - 1 == want["log2_final_degree_plus_1"], key


def test_different_ldts_are_used_for_different_padded_heights():
    """stark.rs:4880-4899: with no forced choice both low-degree tests occur over padded heights 2^0..2^20 (FRI below 2^16)"""
    used = {tvm_b200.derive_domains(8, 2, 1 << h, tvm_b200.LDT_AUTO)["ldt"] for h in range(0, 21)}
    assert used == {tvm_b200.LDT_FRI, tvm_b200.LDT_STIR}
    assert tvm_b200.derive_domains(160, 2, 1 << 15, tvm_b200.LDT_AUTO)["ldt"] == tvm_b200.LDT_FRI
    assert tvm_b200.derive_domains(160, 2, 1 << 16, tvm_b200.LDT_AUTO)["ldt"] == tvm_b200.LDT_STIR

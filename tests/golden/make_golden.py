#!/usr/bin/env python3
"""Regenerates tests/golden/proof_digests.json: Tip5 digests of complete proofs of small synthetic instances and the derived
parameter tables, computed by the CPU oracle (oracle/stark.py).  The reference itself cannot run in this image (no Rust
toolchain), so these fixtures pin the ORACLE against regressions; the GPU path is compared with the same digests.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import stark as S, tip5  # noqa: E402

P = (1 << 64) - (1 << 32) + 1
CASES = [  # (name, security, log2_expansion, ldt, padded_height, seed)
    ("fri_s4_e4_h16", 4, 2, "fri", 16, 101),
    ("fri_s8_e8_h32", 8, 3, "fri", 32, 102),
    ("stir_s6_e4_h256", 6, 2, "stir", 256, 103),
    ("stir_s6_e2_h64", 6, 1, "stir", 64, 104),
]


def rand_bfes(rng, shape):
    return rng.integers(0, P, size=shape, dtype=np.uint64)


def instance(security, log2_exp, ldt, padded_height, seed):
    st = S.Stark(security, log2_exp, ldt)
    d = st.derive(padded_height)
    rng = np.random.default_rng(seed)
    n, h = d["trace_len"], d["num_trace_randomizers"]
    main, mrand = rand_bfes(rng, (379, n)), rand_bfes(rng, (379, h))
    qrand = rand_bfes(rng, (d["num_quotient_randomizer_coefficients"], 3))
    aux_t, aux_r = rand_bfes(rng, (91, n, 3)), rand_bfes(rng, (91, h, 3))
    claim = S.Claim([seed, 2, 3, 4, 5], [7, 8], [9])
    return st, d, claim, main, mrand, (lambda ch: (aux_t, aux_r)), qrand


def main():
    out = {"proofs": {}, "parameters": {}}
    for name, sec, le, ldt, ph, seed in CASES:
        st, d, claim, main_t, mrand, aux, qrand = instance(sec, le, ldt, ph, seed)
        proof, _ = S.prove(st, claim, main_t, mrand, aux, qrand, padded_height=ph)
        assert S.verify(st, claim, proof, check_air=False)
        out["proofs"][name] = {"length": len(proof), "tip5_digest": [int(v) for v in tip5.hash_varlen(proof)]}
    for log2_ph in (10, 16, 20, 22):
        d = S.Stark(160, 2).derive(1 << log2_ph)
        out["parameters"][f"default_2^{log2_ph}"] = {k: d[k] for k in (
            "ldt", "num_trace_randomizers", "trace_len", "ldt_len", "quotient_len", "num_first_round_queries",
            "num_collinearity_checks", "fri_num_rounds")} | ({"stir": d["stir"]} if d["stir"] else {})
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "proof_digests.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out["proofs"], indent=1))


if __name__ == "__main__":
    main()

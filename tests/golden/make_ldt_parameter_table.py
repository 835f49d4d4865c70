#!/usr/bin/env python3
"""Writes tests/golden/ldt_parameter_table.json: the table the reference prints in `print_various_ldt_parameters`
(stark.rs:4901-4958) - for Stark::default() at padded heights 2^8 .. 2^29, FRI and STIR, proven and conjectured regime:
number of rounds, first-round queries, total queries, log2(initial domain length), log2(final degree + 1) - computed by the
ORACLE (oracle/stark.py, oracle/stir.py restate Stark::{fri,stir} and StirParameters::try_into_stir operation for operation).
tests/test_stir_params.py checks tvm_derive_domains (the product's C++ derivation) against it."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import stark as S


def row(ldt, soundness, log2_ph):
    d = S.Stark(160, 2, ldt, soundness).derive(1 << log2_ph)
    if ldt == "fri":
        rounds = d["fri_num_rounds"]
        first = d["num_collinearity_checks"]
        total = first * (rounds + 1)                                   # A indices once + B indices per round (fri.rs LdtStats)
        final_deg_plus_1 = d["fri_last_round_max_degree"] + 1
    else:
        sp = d["stir"]
        rounds = len(sp["round_queries"])
        first = sp["num_first_round_queries"]
        total = sum(a + b for a, b in sp["round_queries"]) + sp["final_num_in_domain_queries"]
        final_deg_plus_1 = sp["final_degree"] + 1
    return dict(num_rounds=rounds, first_round_queries=first, total_queries=total, log2_initial_domain_len=d["ldt_len"].bit_length() - 1,
                log2_final_degree_plus_1=final_deg_plus_1.bit_length() - 1, num_trace_randomizers=d["num_trace_randomizers"],
                trace_len=d["trace_len"])


def main():
    table = {}
    for ldt in ("fri", "stir"):
        for soundness in ("proven", "conjectured"):
            for h in range(8, 30):
                table[f"{ldt}/{soundness}/{h}"] = row(ldt, soundness, h)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ldt_parameter_table.json")
    json.dump(table, open(out, "w"), indent=0, sort_keys=True)
    print("wrote", out, len(table), "rows")


if __name__ == "__main__":
    main()

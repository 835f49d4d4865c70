#!/usr/bin/env python3
"""Writes the input files of examples/prove_tables.c for one of the reference's benchmark workloads:

    python tools/make_workload.py spin_20 /tmp/spin20           # ProgramToBench::spin(20), Stark::default()
    python tools/make_workload.py fib_100 /tmp/fib100 [--security 160] [--ldt auto|fri|stir]
    python tools/make_workload.py verifier_11500 /tmp/ver20      # verifier-shaped program (hashing from memory, u32, X-field, RAM), 2^20
    PROVE_TABLES_REPS=3 ./prove_tables /tmp/spin20              # 149 table columns -> proof, verified incl. the AIR
    python tools/make_workload.py spin_20 /tmp/spin20 --aet     # + the AET arrays: bench.py --workload-dir /tmp/spin20 --from-aet

The tables come from the oracle's VM and table fill (oracle/tracegen.py — bit-identical to the reference on its whole-proof
known-answer tests); this is a TOOL: nothing in the product path or in bench.py imports it.  Only the 149 table columns are
written (the degree-lowering columns are zero: the device fills them); randomness is drawn with numpy from --seed."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload")
    ap.add_argument("out_dir")
    ap.add_argument("--security", type=int, default=160)
    ap.add_argument("--ldt", default="auto", choices=["auto", "fri", "stir"])
    ap.add_argument("--seed", type=int, default=41)
    ap.add_argument("--aet", action="store_true", help="also write the AlgebraicExecutionTrace arrays (aet_*.u64: tvm_aet's fields) "
                    "for bench.py --from-aet / tvm_prove_aet")
    a = ap.parse_args()
    import test_vm_programs as tvp
    program, inp, ram = tvp._workload(a.workload)          # spin_K | fib_N | verifier_N (verifier-shaped mixed program)
    inst = tvp.program_instance(program, inp, a.security, None if a.ldt == "auto" else a.ldt, a.seed, ram=ram)
    os.makedirs(a.out_dir, exist_ok=True)
    main_t = inst["main"].copy()
    main_t[149:] = 0
    for name, arr in (("main", main_t), ("main_rand", inst["main_rand"]), ("aux_rand", inst["aux_rand"]),
                      ("col90", inst["randomizer_column"]), ("quot_rand", inst["quot_rand"])):
        np.ascontiguousarray(arr, dtype="<u8").tofile(os.path.join(a.out_dir, name + ".u64"))
    claim = inst["claim"]
    with open(os.path.join(a.out_dir, "claim.txt"), "w") as f:
        f.write("%d 2 %d %d  %s  %d %s  %d %s\n" % (a.security, {"auto": 0, "fri": 1, "stir": 2}[a.ldt], inst["padded_height"],
                                                  " ".join(str(int(v)) for v in claim.program_digest),
                                                  len(claim.input), " ".join(str(int(v)) for v in claim.input),
                                                  len(claim.output), " ".join(str(int(v)) for v in claim.output)))
    if a.aet:
        from oracle import tracegen as tg
        words = tg.assemble(program)
        arrays = tg.aet_arrays(words, tg.execute(words, inp, (), ram))
        for name, arr in arrays.items():
            dt = "<u4" if name == "instruction_multiplicities" else "<u8"
            np.ascontiguousarray(arr, dtype=dt).tofile(os.path.join(a.out_dir, "aet_%s.%s" % (name, "u32" if dt == "<u4" else "u64")))
    print(a.workload, "padded height", inst["padded_height"], "trace domain", inst["main"].shape[1], "ldt", inst["derived"]["ldt"], "->", a.out_dir)


if __name__ == "__main__":
    main()

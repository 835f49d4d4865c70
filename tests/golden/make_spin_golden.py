#!/usr/bin/env python3
"""Regenerates tests/golden/spin_digests.json: Tip5 digests (Tip5::hash(&proof)) of the ORACLE's proofs of the reference's
benchmark workloads — ProgramToBench::spin(k) (triton-dev-util/src/lib.rs:49-75) and Fibonacci(100) (benches/prove_fib.rs) —
under Stark::default() — at 2^16 that means STIR.
The oracle is bit-identical to the reference on the reference's own whole-proof known-answer tests (tests/test_golden.py);
these fixtures extend that anchor to the BASELINE configuration "padded height 2^16" without re-running the oracle
(≈ 5 minutes) in every test run.  The GPU test compares tvm_prove's proof of the same instance with the digest.

    python tests/golden/make_spin_golden.py [workload ...]        (default: fib_100 spin_13 spin_16 spin_18; also verifier_700)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "triton-vm_b200", "py")]
from oracle import corc, fast, reference_prover as RP, stark as S  # noqa: E402
import test_vm_programs as tvp  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "spin_digests.json")


def main():
    names = sys.argv[1:] or ["fib_100", "spin_13", "spin_16", "spin_18"]
    out = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name in names:
        program, inp, ram = tvp._workload(name)
        inst = tvp.program_instance(program, inp, 160, None, ram=ram)
        claim, main_t, rcol, arand = inst["claim"], inst["main"], inst["randomizer_column"], inst["aux_rand"]
        t = time.time()
        # oracle/fast.py = the same proof words as oracle/stark.py's prove (tests/test_oracle_fast.py), array stages in C
        prove = (lambda *a, **k: (fast.prove(*a, **k), None)) if os.environ.get("ORACLE_FAST", "1") != "0" else S.prove
        proof, _ = prove(inst["stark"], claim, main_t, inst["main_rand"],
                         lambda ch: (corc.aux_extend(main_t, np.asarray(ch, dtype=np.uint64).reshape(63, 3), rcol), arand),
                         inst["quot_rand"], padded_height=inst["padded_height"])
        import tvm_b200
        assert tvm_b200.verify((claim.program_digest, claim.input, claim.output), proof, 160, 2) == (True, "")
        out[name] = {"ldt": inst["derived"]["ldt"], "proof_words": len(proof), "tip5_digest": RP.proof_digest(proof),
                            "oracle_seconds": round(time.time() - t)}
        print(name, out[name], flush=True)
        json.dump(out, open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main()

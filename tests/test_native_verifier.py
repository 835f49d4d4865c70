"""tvm_verify — Stark::verify in the library (csrc/verifier.cu, host code: needs no GPU, so these tests run everywhere).
It must accept what the oracle's verifier accepts — first of all the two proofs whose Tip5 digests are the reference's own
known-answer values (tests/test_golden.py), i.e. proofs bit-identical to what the reference produces — and reject what it
rejects: every tampered word, a wrong claim, wrong parameters."""
import importlib.util
import os

import numpy as np
import pytest

import tvm_b200
from oracle import reference_prover as RP, stark as S
import test_golden as tgold
from test_halt_program import _instance as halt_instance, halt_tables, N as HALT_N

_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mg)
_PROOFS = {}


def _c(claim):
    return (claim.program_digest, claim.input, claim.output)


def reference_proof(which):
    if which not in _PROOFS:
        inst = tgold.reference_instance() if which == "default" else tgold.every_instruction_instance()
        _PROOFS[which] = (inst, RP.prove(inst))
    return _PROOFS[which]


@pytest.mark.parametrize("which,security", [("default", 160), ("every_instruction", 32)])
def test_accepts_the_references_known_answer_proofs(which, security):
    inst, proof = reference_proof(which)
    assert RP.proof_digest(proof) == (tgold.REFERENCE_PROOF_DIGEST if which == "default" else tgold.EVERY_INSTRUCTION_PROOF_DIGEST)
    ok, why = tvm_b200.verify(_c(inst["claim"]), proof, security_level=security, log2_expansion=2)
    assert ok, why


def test_rejects_tampering_wrong_claims_and_wrong_parameters():
    inst, proof = reference_proof("default")
    claim = inst["claim"]
    rng = np.random.default_rng(1)
    reasons = set()
    for pos in [0, 1, 2, 3, 9, 40, 1500, 3000, len(proof) // 2, len(proof) - 1] + [int(v) for v in rng.integers(0, len(proof), 12)]:
        bad = list(proof)
        bad[pos] = (bad[pos] + 1) % tvm_b200.P
        ok, why = tvm_b200.verify(_c(claim), bad, 160, 2)
        assert not ok and why, pos
        with pytest.raises(Exception):
            S.verify(inst["stark"], claim, bad, check_air=True)
        reasons.add(why.split(":")[0])
    assert {"ProofDecodingError", "VerificationError", "LdtVerificationError"} <= reasons
    assert tvm_b200.verify(_c(claim), proof[:-1], 160, 2)[0] is False                     # truncated
    assert tvm_b200.verify((claim.program_digest, claim.input, [7]), proof, 160, 2) == (False, "VerificationError: OutOfDomainQuotientValueMismatch")
    assert tvm_b200.verify((claim.program_digest[::-1], claim.input, claim.output), proof, 160, 2)[0] is False
    assert tvm_b200.verify(_c(claim), proof, 80, 2)[0] is False                            # other security level
    assert tvm_b200.verify(_c(claim), proof, 160, 3)[0] is False                           # other expansion factor
    assert tvm_b200.verify(_c(claim), proof, 160, 2, ldt_choice=tvm_b200.LDT_STIR)[0] is False


@pytest.mark.parametrize("case", mg.CASES, ids=lambda c: c[0])
def test_accepts_synthetic_fri_and_stir_proofs_without_the_air_check(case):
    name, sec, le, ldt, ph, seed = case
    st, d, claim, main, mrand, aux, qrand = mg.instance(sec, le, ldt, ph, seed)
    proof, _ = S.prove(st, claim, main, mrand, aux, qrand, padded_height=ph)
    choice = tvm_b200.LDT_STIR if ldt == "stir" else tvm_b200.LDT_FRI
    assert tvm_b200.verify(_c(claim), proof, sec, le, ldt_choice=choice, skip_air_check=True) == (True, "")
    ok, why = tvm_b200.verify(_c(claim), proof, sec, le, ldt_choice=choice)             # random tables do not satisfy the AIR
    assert (ok, why) == (False, "VerificationError: OutOfDomainQuotientValueMismatch")
    for pos in (len(proof) // 3, len(proof) - 40):
        bad = list(proof)
        bad[pos] = (bad[pos] + 1) % tvm_b200.P
        assert tvm_b200.verify(_c(claim), bad, sec, le, ldt_choice=choice, skip_air_check=True)[0] is False


def test_accepts_a_stir_proof_of_a_real_program_with_the_air_check():
    st, claim, main, mrand, aux_provider, qrand = halt_instance(halt_tables(HALT_N), 8, "stir")
    proof, _ = S.prove(st, claim, main, mrand, aux_provider, qrand, padded_height=HALT_N)
    assert tvm_b200.verify(_c(claim), proof, 8, 2, ldt_choice=tvm_b200.LDT_STIR) == (True, "")
    assert tvm_b200.verify(_c(claim), proof, 8, 2, ldt_choice=tvm_b200.LDT_FRI)[0] is False


def test_argument_errors_and_padded_height():
    with pytest.raises(tvm_b200.TvmError):
        tvm_b200.verify(([1, 2, 3, 4, 5], [], []), [], 160, 2)
    inst, proof = reference_proof("default")
    assert tvm_b200.proof_padded_height(proof) == inst["padded_height"] == 256          # Proof::padded_height, proof.rs:37-56
    with pytest.raises(tvm_b200.TvmError):
        tvm_b200.proof_padded_height(proof[:100])
    l = tvm_b200.lib()                                   # device entry points refuse a missing context before touching CUDA
    assert l.tvm_aux_extend(None, None, 8, None, None, None) == -1 and l.tvm_fill_derived_main_columns(None, None, 8) == -1


def test_untrusted_input_never_crashes_and_is_never_accepted():
    # the reference's `decoding_arbitrary_proof_data_does_not_panic` (proof.rs:192-197), extended to the whole verifier:
    # random words, corrupted lengths / counts / discriminants, truncation, extension, absurd padded heights
    rng = np.random.default_rng(7)
    cases = []
    for name, sec, le, ldt, ph, seed in (mg.CASES[0], mg.CASES[3]):
        st, d, claim, main, mrand, aux, qrand = mg.instance(sec, le, ldt, ph, seed)
        proof, _ = S.prove(st, claim, main, mrand, aux, qrand, padded_height=ph)
        cases.append((claim, np.array(proof, dtype=np.uint64), sec, le, tvm_b200.LDT_STIR if ldt == "stir" else tvm_b200.LDT_FRI))
    outcomes = set()
    for k in range(6000):
        claim, proof, sec, le, choice = cases[k % 2]
        p = proof.copy()
        kind = k % 6
        if kind == 0:
            p[rng.integers(0, len(p))] = rng.integers(0, tvm_b200.P, dtype=np.uint64)
        elif kind == 1:
            p[rng.integers(0, len(p))] = rng.integers(0, 70000)
        elif kind == 2:
            p = p[:rng.integers(1, len(p))] if k % 4 else np.concatenate([p, rng.integers(0, tvm_b200.P, 17, dtype=np.uint64)])
        elif kind == 3:
            p[rng.integers(0, 60)] = rng.integers(0, 40)
        elif kind == 4:
            p[rng.integers(0, len(p))] = np.uint64(2 ** 64 - 1 - int(rng.integers(0, 2 ** 33)))
        else:
            p = rng.integers(0, tvm_b200.P, int(rng.integers(1, 3000)), dtype=np.uint64)
        ok, why = tvm_b200.verify(_c(claim), p, sec, le, ldt_choice=choice, skip_air_check=True)
        assert not ok or np.array_equal(p, proof)
        outcomes.add(why.split(":")[0])
    assert {"ProofDecodingError", "VerificationError", "LdtVerificationError", "UnexpectedItem"} <= outcomes
    claim, proof, sec, le, choice = cases[0]
    for log2_ph in list(range(5, 34)) + [63, 2 ** 32, tvm_b200.P - 1]:                # absurd heights cost nothing
        p = proof.copy()
        p[4] = log2_ph
        assert tvm_b200.verify(_c(claim), p, sec, le, ldt_choice=choice, skip_air_check=True)[0] is False


def test_plain_c_client(tmp_path):
    # include/tvm_b200.h is a C header: compile examples/verify_proof.c with a C99 compiler and let it judge the reference's
    # known-answer proof and a corrupted copy
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "triton-vm_b200", "lib")
    exe = str(tmp_path / "verify_proof")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(root, "include"),
                    os.path.join(root, "examples", "verify_proof.c"), "-L" + libdir, "-ltvm_b200", "-Wl,-rpath," + libdir, "-o", exe],
                   check=True, env=dict(os.environ, CC="gcc"))
    inst, proof = reference_proof("default")
    claim = inst["claim"]

    def write(path, words):
        head = [160, 2, 0, claim.version] + list(claim.program_digest) + [len(claim.input)] + list(claim.input) + \
               [len(claim.output)] + list(claim.output) + [len(words)] + [int(w) for w in words]
        path.write_text(" ".join(str(int(v)) for v in head))
    write(tmp_path / "good.txt", proof)
    bad = list(proof)
    bad[len(bad) // 2] ^= 1
    write(tmp_path / "bad.txt", bad)
    good = subprocess.run([exe, str(tmp_path / "good.txt")], capture_output=True, text=True)
    assert good.returncode == 0 and "accepted (padded height 256" in good.stdout, good.stderr
    rejected = subprocess.run([exe, str(tmp_path / "bad.txt")], capture_output=True, text=True)
    assert rejected.returncode == 1 and "rejected:" in rejected.stderr


def test_batch_verification_on_host_threads():
    inst, proof = reference_proof("default")
    claim = _c(inst["claim"])
    bad = list(proof)
    bad[777] ^= 1
    claims = [claim] * 6 + [(claim[0], claim[1], [3])]
    proofs = [proof, bad, proof, proof, bad, proof, proof]
    assert tvm_b200.verify_batch(claims, proofs, 160, 2, num_threads=4) == [True, False, True, True, False, True, False]
    assert tvm_b200.verify_batch(claims[:1], proofs[:1], 160, 2) == [True] and tvm_b200.verify_batch([], [], 160, 2) == []


def _write_prove_tables_dir(tmp_path):
    from conftest import rand_bfes
    T, digest, main = halt_tables(HALT_N)
    st = S.Stark(8, 2, "fri")
    d = st.derive(HALT_N)
    n, h = d["trace_len"], d["num_trace_randomizers"]
    assert n == HALT_N
    rng = np.random.default_rng(77)
    base = main.copy()
    base[149:] = 0                                            # the example fills the degree-lowering columns on the device
    for name, arr in (("main", base), ("main_rand", rand_bfes(rng, (379, h))), ("aux_rand", rand_bfes(rng, (91, h, 3))),
                      ("col90", rand_bfes(rng, (n, 3))), ("quot_rand", rand_bfes(rng, (d["num_quotient_randomizer_coefficients"], 3)))):
        np.ascontiguousarray(arr, dtype="<u8").tofile(str(tmp_path / (name + ".u64")))
    (tmp_path / "claim.txt").write_text("8 2 1 %d  %s  0  0\n" % (HALT_N, " ".join(str(int(v)) for v in digest)))
    return digest


def _build_example(tmp_path, name):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "triton-vm_b200", "lib")
    exe = str(tmp_path / name)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(root, "include"),
                    os.path.join(root, "examples", name + ".c"), "-L" + libdir, "-ltvm_b200", "-Wl,-rpath," + libdir, "-o", exe], check=True)
    return exe


def test_plain_c_prover_client_builds_and_fails_loudly_without_a_gpu(tmp_path):
    import subprocess
    exe = _build_example(tmp_path, "prove_tables")
    _write_prove_tables_dir(tmp_path)
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    if tvm_b200.lib().tvm_device_count() == 0:
        assert r.returncode == 3 and "no CPU fallback" in r.stderr           # tvm_ctx_create refuses: nothing is computed on the host
    else:
        assert r.returncode == 0 and "verified" in r.stdout, r.stderr


def _write_aet_files(tmp_path):
    """the AET of `halt` next to the files of _write_prove_tables_dir (tools/make_workload.py --aet writes the same)"""
    from oracle import tracegen as tg
    words = [tg.OP_HALT]
    for name, arr in tg.aet_arrays(words, tg.execute(words)).items():
        u32 = name == "instruction_multiplicities"
        np.ascontiguousarray(arr, dtype="<u4" if u32 else "<u8").tofile(str(tmp_path / ("aet_%s.%s" % (name, "u32" if u32 else "u64"))))


def test_plain_c_aet_client_builds_and_fails_loudly_without_a_gpu(tmp_path):
    import subprocess
    exe = _build_example(tmp_path, "prove_aet")
    _write_prove_tables_dir(tmp_path)
    _write_aet_files(tmp_path)
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    if tvm_b200.lib().tvm_device_count() == 0:
        assert r.returncode == 3 and "no CPU fallback" in r.stderr
    else:
        assert r.returncode == 0 and "verified" in r.stdout, r.stderr


def test_different_ldts_and_proximity_regimes_are_mutually_incompatible():
    """different_ldts_are_mutually_incompatible / different_proximty_regimes_are_mutually_incompatible (stark.rs:4840-4876) on the
    program `halt`: a proof verifies under exactly the (low-degree test, proximity regime) it was made for"""
    from oracle import fast
    tables = halt_tables(HALT_N)
    proofs = {}
    for ldt in ("fri", "stir"):
        for regime in ("proven", "conjectured"):
            st, claim, main, mrand, aux_provider, qrand = halt_instance(tables, 8, ldt)
            st = S.Stark(8, 2, ldt, regime)
            d = st.derive(HALT_N)
            rng = np.random.default_rng(5)
            h = d["num_trace_randomizers"]
            mrand = rng.integers(0, tvm_b200.P, size=(379, h), dtype=np.uint64)
            arand = rng.integers(0, tvm_b200.P, size=(91, h, 3), dtype=np.uint64)
            qrand = rng.integers(0, tvm_b200.P, size=(d["num_quotient_randomizer_coefficients"], 3), dtype=np.uint64)
            proofs[(ldt, regime)] = (claim, fast.prove(st, claim, main, mrand, lambda ch: (aux_provider(ch)[0], arand), qrand, padded_height=HALT_N))
    choice = {"fri": tvm_b200.LDT_FRI, "stir": tvm_b200.LDT_STIR}
    for (ldt, regime), (claim, proof) in proofs.items():
        for v_ldt in ("fri", "stir"):
            for v_regime in ("proven", "conjectured"):
                ok, why = tvm_b200.verify(_c(claim), proof, 8, 2, ldt_choice=choice[v_ldt], conjectured=v_regime == "conjectured")
                assert ok == ((v_ldt, v_regime) == (ldt, regime)), (ldt, regime, v_ldt, v_regime, why)

"""An AIR-SATISFYING instance: the program `halt`.  oracle/tracegen.py builds its master tables (main table from the
reference's fill/pad rules, auxiliary table by solving the AIR's own initial/transition constraints); every one of the
604 constraints vanishes where it must, and a proof of it passes the verifier INCLUDING the out-of-domain AIR /
quotient identity (stark.rs:1469-1540) that synthetic tables can never satisfy."""
import numpy as np
import pytest

from conftest import rand_bfes
from oracle import field as F, stark as S, tip5, tracegen as tg

N = 256          # padded height of `halt`: the lookup table's 256 rows (aet.rs:101)
_TABLES = {}


def halt_tables(n):
    if n not in _TABLES:
        T, digest = tg.halt_main_table(n)
        _TABLES[n] = (T, digest, np.array(T.tolist(), dtype=np.uint64))
    return _TABLES[n]


@pytest.fixture(scope="module")
def halt():
    return halt_tables(N)


def _instance(_halt, security, ldt, seed=9):
    st = S.Stark(security, 2, ldt)
    d = st.derive(N)
    T, digest, main = halt_tables(d["trace_len"])     # the trace domain exceeds the padded height at high security
    h = d["num_trace_randomizers"]
    rng = np.random.default_rng(seed)
    mrand, arand = rand_bfes(rng, (379, h)), rand_bfes(rng, (91, h, 3))
    qrand = rand_bfes(rng, (d["num_quotient_randomizer_coefficients"], 3))

    def aux_provider(ch):
        ch63 = [tuple(int(v) for v in row) for row in np.asarray(ch, dtype=np.uint64).reshape(63, 3)]
        A, _ = tg.extend_by_solving(T, ch63)
        return np.array(A, dtype=np.uint64), arand

    return st, S.Claim(digest, [], []), main, mrand, aux_provider, qrand


def test_program_digest_and_air_satisfaction(halt):
    T, digest, _ = halt
    assert digest == [int(v) for v in tip5.hash_varlen([tg.OP_HALT])]        # Program::hash of [halt]
    rng = np.random.default_rng(5)
    sampled = [tuple(int(v) for v in rng.integers(0, F.P, 3, dtype=np.uint64)) for _ in range(59)]
    challenges = S.derive_challenges(sampled, S.Claim(digest, [], []))
    A, _ = tg.extend_by_solving(T, challenges)
    assert tg.failing_constraints(T, A, challenges) == []
    # and the check is not vacuous: a single flipped table entry is caught
    T2 = T.copy()
    T2[tg.MAIN["processor"].ST3, 7] = 5
    assert tg.failing_constraints(T2, A, challenges) != []


@pytest.mark.parametrize("ldt", ["fri", "stir"])
def test_oracle_proof_of_halt_verifies_with_air_check(halt, ldt):
    st, claim, main, mrand, aux_provider, qrand = _instance(halt, 8, ldt)
    proof, _ = S.prove(st, claim, main, mrand, aux_provider, qrand, padded_height=N)
    assert S.verify(st, claim, proof, check_air=True)
    with pytest.raises(ValueError):
        S.verify(st, S.Claim([1, 2, 3, 4, 5], [], []), proof, check_air=True)


@pytest.mark.gpu
@pytest.mark.parametrize("security,ldt", [(160, "fri"), (8, "stir")])   # Stark::default() security: trace domain 512
def test_gpu_proof_of_halt_verifies_with_air_check(backend, halt, security, ldt):
    import tvm_b200
    st, claim, main, mrand, aux_provider, qrand = _instance(halt, security, ldt)
    got = backend.prove((claim.program_digest, claim.input, claim.output), main, mrand, aux_provider, qrand,
                        security_level=security, log2_expansion=2, padded_height=N,
                        ldt_choice=tvm_b200.LDT_STIR if ldt == "stir" else tvm_b200.LDT_FRI)
    got = [int(v) for v in got]
    assert S.verify(st, claim, got, check_air=True)          # a complete, valid STARK proof of `halt`
    want, _ = S.prove(st, claim, main, mrand, aux_provider, qrand, padded_height=N)
    assert got == want

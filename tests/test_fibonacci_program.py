"""AIR-SATISFYING instances beyond `halt`: oracle/tracegen.py's small VM runs real programs (the reference's Fibonacci
example, triton-vm/examples + vm.rs tests, and a branchy arithmetic program), builds all nine tables from the recorded
execution (processor, op stack, jump stack, program, hash, cascade, lookup; RAM and u32 stay empty), and the auxiliary
table is solved from the AIR itself.  All 604 constraints vanish, the cross-table arguments close (terminal constraints:
permutation / evaluation / lookup arguments against the claim's input and output), and proofs verify INCLUDING the
out-of-domain AIR identity (stark.rs:1469-1540)."""
import numpy as np
import pytest

from conftest import rand_bfes
from oracle import field as F, stark as S, tracegen as tg

FIBONACCI = """
    push 0 push 1 read_io 1
    dup 0 skiz call fib_loop
    pop 1 write_io 1 halt
  fib_loop:
    push -1 add swap 2 dup 1 add swap 1 swap 2 dup 0 skiz recurse return
"""

# eq / mul / assert / nop, a taken and a not-taken skiz, a skipped two-word instruction, nested calls
BRANCHY = """
    read_io 2 dup 1 dup 1 eq skiz push 99 nop
    mul dup 0 push 35 eq assert
    call outer write_io 1 push 0 skiz halt halt
  outer:
    call inner push 3 mul return
  inner:
    push 4 add return
"""

_CACHE = {}


def tables(program, public_input, n=None, secret_input=()):
    key = (program, tuple(public_input), n, tuple(secret_input))
    if key not in _CACHE:
        words = tg.assemble(program)
        ph = tg.padded_height(words, public_input, secret_input)
        T, digest, out = tg.main_table(words, public_input, n or ph, secret_input)
        _CACHE[key] = (T, digest, out, ph, np.array(T, dtype=np.uint64))
    return _CACHE[key]


def test_vm_runs_fibonacci():
    words = tg.assemble(FIBONACCI)
    fib = [1, 1, 2, 3, 5, 8, 13, 21, 34, 55, 89]
    for k in range(11):
        rows, _, mult, out, _ = tg.run(words, [k])
        assert out == [fib[k]]
        assert sum(mult) == len(rows)
    with pytest.raises(IndexError):
        tg.run(words, [])                                   # reading past the public input


def _challenges(digest, inp, out, seed=5):
    rng = np.random.default_rng(seed)
    sampled = [tuple(int(v) for v in rng.integers(0, F.P, 3, dtype=np.uint64)) for _ in range(59)]
    return S.derive_challenges(sampled, S.Claim(digest, list(inp), list(out)))


@pytest.mark.parametrize("program,inp,want_out", [(FIBONACCI, [7], [21]), (FIBONACCI, [0], [1]), (BRANCHY, [5, 7], [(35 + 4) * 3])])
def test_every_constraint_vanishes(program, inp, want_out):
    T, digest, out, ph, _ = tables(program, inp)
    assert out == want_out and ph == 512                     # dominated by the hash table's cascade lookups
    ch = _challenges(digest, inp, out)
    A, _ = tg.extend_by_solving(T, ch)
    assert tg.failing_constraints(T, A, ch) == []
    # a wrong claimed output breaks exactly the cross-table link to the claim
    bad = tg.failing_constraints(T, A, _challenges(digest, inp, [out[0] + 1]))
    assert bad and all(kind == "term" for kind, *_ in bad)


def _instance(program, inp, security, ldt, seed=11):
    st = S.Stark(security, 2, ldt)
    ph = tables(program, inp)[3]
    d = st.derive(ph)
    T, digest, out, _, main = tables(program, inp, d["trace_len"])
    h = d["num_trace_randomizers"]
    rng = np.random.default_rng(seed)
    mrand, arand = rand_bfes(rng, (379, h)), rand_bfes(rng, (91, h, 3))
    qrand = rand_bfes(rng, (d["num_quotient_randomizer_coefficients"], 3))

    def aux_provider(ch):
        ch63 = [tuple(int(v) for v in row) for row in np.asarray(ch, dtype=np.uint64).reshape(63, 3)]
        return np.array(tg.extend_by_solving(T, ch63)[0], dtype=np.uint64), arand

    return st, S.Claim(digest, list(inp), list(out)), main, mrand, aux_provider, qrand, ph


def test_oracle_proof_of_fibonacci_verifies_with_air_check():
    st, claim, main, mrand, aux_provider, qrand, ph = _instance(FIBONACCI, [7], 8, "fri")
    proof, _ = S.prove(st, claim, main, mrand, aux_provider, qrand, padded_height=ph)
    assert S.verify(st, claim, proof, check_air=True)
    with pytest.raises(ValueError):                          # fib(7) is not 22
        S.verify(st, S.Claim(claim.program_digest, [7], [22]), proof, check_air=True)


@pytest.mark.gpu
@pytest.mark.parametrize("program,inp,ldt", [(FIBONACCI, [7], "fri"), (BRANCHY, [5, 7], "stir")])
def test_gpu_proof_of_program_verifies_with_air_check(backend, program, inp, ldt):
    import tvm_b200
    st, claim, main, mrand, aux_provider, qrand, ph = _instance(program, inp, 8, ldt)
    got = backend.prove((claim.program_digest, claim.input, claim.output), main, mrand, aux_provider, qrand,
                        security_level=8, log2_expansion=2, padded_height=ph,
                        ldt_choice=tvm_b200.LDT_STIR if ldt == "stir" else tvm_b200.LDT_FRI)
    got = [int(v) for v in got]
    assert S.verify(st, claim, got, check_air=True)
    want, _ = S.prove(st, claim, main, mrand, aux_provider, qrand, padded_height=ph)
    assert got == want

"""Device-side auxiliary-table extension (tvm_aux_extend, csrc/aux_extend.cu: per-column scan of the row maps derived
from the AIR) against the CPU checker that runs the same generated rules as the reference's sequential loop
(oracle/c/aux_extend.c; itself pinned to the AIR-solving restatement in tests/test_aux_extend.py), and inside a complete
prove(): the host callback only forwards the device's table."""
import numpy as np
import pytest

from conftest import rand_bfes
from oracle import corc, field as F, stark as S, tracegen as tg
from test_aux_extend import semi_valid_table, _challenges
from test_fibonacci_program import FIBONACCI, tables

pytestmark = pytest.mark.gpu
P = F.P


@pytest.mark.parametrize("n,seed,npad", [(256, 11, 9), (1024, 12, 100), (4096, 13, 0)])   # 1, 4 and 16 scan chunks
def test_device_extension_equals_cpu_rules_on_every_instruction(backend, n, seed, npad):
    T = np.array(semi_valid_table(n, seed, npad).tolist(), dtype=np.uint64)
    ch = _challenges(seed)
    rc = rand_bfes(np.random.default_rng(seed), (n, 3))
    want = corc.aux_extend(T, ch, rc)
    got = backend.aux_extend(T, ch, rc)
    bad = [q for q in range(91) if not np.array_equal(got[q], want[q])]
    where = {q: (int(np.nonzero((got[q] != want[q]).any(axis=1))[0][0]), int((got[q] != want[q]).any(axis=1).sum())) for q in bad[:8]}
    assert bad == [], f"columns {bad} differ; (first bad row, #bad rows) = {where}"
    assert np.array_equal(backend.aux_extend(T, ch)[90], np.zeros((n, 3), dtype=np.uint64))      # no randomizer column given


def _prove_with_device_extension(backend, program, inp, security, ldt):
    import tvm_b200
    st = S.Stark(security, 2, ldt)
    ph = tables(program, inp)[3]
    d = st.derive(ph)
    T, digest, out, _, main = tables(program, inp, d["trace_len"])
    n, h = d["trace_len"], d["num_trace_randomizers"]
    rng = np.random.default_rng(21)
    mrand, arand, rcol = rand_bfes(rng, (379, h)), rand_bfes(rng, (91, h, 3)), rand_bfes(rng, (n, 3))
    qrand = rand_bfes(rng, (d["num_quotient_randomizer_coefficients"], 3))
    claim = S.Claim(digest, list(inp), list(out))
    calls = []

    def device_extend(ch):
        calls.append(1)
        return backend.aux_extend(main, np.asarray(ch, dtype=np.uint64).reshape(63, 3), rcol), arand

    def cpu_extend(ch):
        return corc.aux_extend(main, [tuple(int(v) for v in row) for row in np.asarray(ch, dtype=np.uint64).reshape(63, 3)], rcol), arand

    got = backend.prove((claim.program_digest, claim.input, claim.output), main, mrand, device_extend, qrand,
                        security_level=security, log2_expansion=2, padded_height=ph,
                        ldt_choice=tvm_b200.LDT_STIR if ldt == "stir" else tvm_b200.LDT_FRI)
    got = [int(v) for v in got]
    assert calls == [1]
    assert S.verify(st, claim, got, check_air=True)              # the device-extended table satisfies the AIR
    want, _ = S.prove(st, claim, main, mrand, cpu_extend, qrand, padded_height=ph)
    assert got == want


def test_prove_halt_with_device_side_extension(backend):
    _prove_with_device_extension(backend, "halt", [], 8, "fri")


def test_prove_fibonacci_with_device_side_extension(backend):
    _prove_with_device_extension(backend, FIBONACCI, [7], 8, "stir")

"""MasterMainTable::new + pad from the AET (master_table.rs:881-974): the stage bodies and the orchestration of
triton-vm_b200/csrc/fill/main_fill.cuh, run by the sequential host executor of tests/host/main_fill_host.cu (test
infrastructure; the library instantiates the same header with the CUDA executor only), against the oracle's table fill —
which is pinned by the reference's whole-proof known-answer tests.  The GPU run of the same checks is
tests/test_z_gpu_main_fill.py."""
import ctypes
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "triton-vm_b200"), os.path.join(ROOT, "triton-vm_b200", "py"), os.path.dirname(os.path.abspath(__file__))]
from oracle import tracegen as tg   # noqa: E402
import tvm_b200                     # noqa: E402
import test_vm_programs as tvp      # noqa: E402

P = tg.P
SRC = os.path.join(ROOT, "tests", "host", "main_fill_host.cu")
SO = os.path.join(ROOT, "tests", "host", "_build", "libmain_fill_host.so")
HDR = os.path.join(ROOT, "triton-vm_b200", "csrc", "fill", "main_fill.cuh")


@pytest.fixture(scope="module")
def host():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    stale = not os.path.exists(SO) or any(os.path.getmtime(f) > os.path.getmtime(SO) for f in (SRC, HDR))
    if stale:
        if not os.path.exists(nvcc):
            pytest.skip("nvcc not available to build the host harness")
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.run([nvcc, "-O2", "-std=c++17", "--extended-lambda", "--expt-relaxed-constexpr", "-shared", "-Xcompiler", "-fPIC",
                        SRC, "-o", SO], check=True, capture_output=True)
    return ctypes.CDLL(SO)


def host_bezout(lib, roots, direct_log):
    m = len(roots)
    r = np.array(roots, dtype=np.uint64)
    a, b, st = np.zeros(m, dtype=np.uint64), np.zeros(m, dtype=np.uint64), np.zeros(2, dtype=np.uint64)
    vp = ctypes.c_void_p
    rc = lib.fill_host_bezout(r.ctypes.data_as(vp), ctypes.c_uint64(m), a.ctypes.data_as(vp), b.ctypes.data_as(vp), ctypes.c_uint(direct_log),
                              st.ctypes.data_as(vp))
    assert rc == 0
    return [int(x) for x in a], [int(x) for x in b], [int(x) for x in st]


def host_main_table(lib, arrays, n, direct_log=5):
    s, keep = tvm_b200.aet_struct(arrays)
    out = np.zeros((149, n), dtype=np.uint64)
    lengths = np.zeros(9, dtype=np.uint64)
    err = ctypes.create_string_buffer(256)
    rc = lib.fill_host_main_table(ctypes.byref(s), ctypes.c_uint64(n), out.ctypes.data_as(ctypes.c_void_p), lengths.ctypes.data_as(ctypes.c_void_p),
                                  ctypes.c_uint(direct_log), err, ctypes.c_size_t(256))
    del keep
    if rc:
        raise ValueError(err.value.decode())
    return out, [int(v) for v in lengths]


@pytest.mark.parametrize("m", [1, 2, 3, 5, 8, 9, 31, 32, 33, 100])
@pytest.mark.parametrize("direct_log", [0, 2, 5])
def test_bezout_coefficients_match_the_oracle(host, m, direct_log):
    """ram.rs:162-214 — direct_log 0: every tree level through transforms (down to size 2); 5: the library's setting"""
    rng = np.random.default_rng(1000 * m + direct_log)
    roots = [int(x) for x in rng.integers(0, P, size=m, dtype=np.uint64)]
    if m > 3:
        roots[:3] = [0, 1, P - 1]
    roots = list(dict.fromkeys(roots))
    a, b, _ = host_bezout(host, roots, direct_log)
    want_a, want_b = tg.bezout_coefficients(roots)
    assert a == want_a and b == want_b


def _ev(poly, z):
    acc = 0
    for c in reversed(poly):
        acc = (acc * z + c) % P
    return acc


@pytest.mark.parametrize("m,direct_log", [(1000, 5), (4097, 3), (6000, 5)])
def test_bezout_identity_on_larger_root_sets(host, m, direct_log):
    """a rp + b rp' = 1 with deg a < m - 1, deg b < m determines a and b; checked at points outside the roots"""
    rng = np.random.default_rng(m)
    roots = list(dict.fromkeys(int(x) for x in rng.integers(0, P, size=m, dtype=np.uint64)))
    roots[:4] = [7, 8, 9, 10]                       # consecutive pointers, as arrays in RAM are
    roots = list(dict.fromkeys(roots))
    a, b, stats = host_bezout(host, roots, direct_log)
    assert a[-1] == 0
    for z in (3, 12345678901234567, P - 2):
        rp, fd = 1, 0
        for r in roots:
            fd = (fd * (z - r) + rp) % P
            rp = rp * (z - r) % P
        assert (_ev(a, z) * rp + _ev(b, z) * fd) % P == 1
    assert stats[1] > 0                             # the transform paths ran


PROGRAMS = ["halt", "fib_100", "spin_5", "verifier_3", "verifier_40", "u32_mix"]


def _program(name):
    if name == "halt":
        return [tg.OP_HALT], [], None
    if name == "u32_mix":
        src = """
            push 0 push 0 lt pop 1            push 5 push 9 lt pop 1           push 9 push 5 lt pop 1   push 7 push 7 lt pop 1
            push 4294967295 push 1 and pop 1  push 0 push 12 xor pop 1         push 0 log_2_floor pop 1 hint_skip: push 1 log_2_floor pop 1
            push 1000 log_2_floor pop 1       push 0 push 3 pow pop 1          push 13 push 18446744069414584320 pow pop 1
            push 0 push 0 pow pop 1           push 0 pop_count pop 1           push 4294967295 pop_count pop 1
            push 17 split pop 2               push 18446744069414584320 split pop 2   push 3 push 100 div_mod pop 2
            push 5 push 9 lt pop 1            halt"""
        src = src.replace("push 0 log_2_floor pop 1 hint_skip:", "")       # log_2_floor of 0 is a VM error
        return tg.assemble(src), [], None
    src, inp, ram = tvp._workload(name)
    return tg.assemble(src), inp, ram


@pytest.mark.parametrize("name", PROGRAMS)
@pytest.mark.parametrize("extra", [0, 1])
def test_tables_from_the_aet_match_the_oracle(host, name, extra):
    program, inp, ram = _program(name)
    ex = tg.execute(list(program), inp, (), ram)
    n = max(256, tg.padded_height(list(program), inp, (), ram)) << extra      # also a trace domain above the padded height
    want, _, _ = tg.main_table(list(program), inp, n, (), ram)
    arrays = tg.aet_arrays(list(program), ex)
    for direct_log in ((5, 1) if name.startswith("verifier") else (5,)):
        got, lengths = host_main_table(host, arrays, n, direct_log)
        bad = [c for c in range(149) if not np.array_equal(got[c], want[c])]
        assert not bad, [tg.column_name(True, c) for c in bad[:8]]
    heights = tg.table_heights(list(program), ex)
    assert lengths == [heights[k] for k in ("program", "processor", "op_stack", "ram", "jump_stack", "hash", "cascade", "lookup", "u32")]


def test_fill_rejects_a_table_height_below_a_table(host):
    program, inp, ram = _program("fib_100")
    ex = tg.execute(list(program), inp, (), ram)
    arrays = tg.aet_arrays(list(program), ex)
    with pytest.raises(ValueError):
        host_main_table(host, arrays, 256)


def test_u32_sections_match_the_oracle_on_random_entries(host):
    """u32.rs:193-290 on operands no test program produces: every u32 instruction with random and edge-case operands (0, 1,
    2^32 - 1, powers of two; `pow` with a full-width base).  The fill does not look at the consistency of the AET, so the
    entries are planted into the AET of `halt`."""
    rng = np.random.default_rng(99)
    ops = ["split", "lt", "and", "log_2_floor", "pow", "pop_count"]
    edge = [0, 1, 2, 3, (1 << 31), (1 << 32) - 1, (1 << 16), 0x55555555]
    entries = []
    for op in ops:
        pairs = [(a, b) for a in edge for b in edge] + [(int(a), int(b)) for a, b in rng.integers(0, 1 << 32, size=(40, 2))]
        for lhs, rhs in pairs:
            if op == "log_2_floor" and lhs == 0:
                continue                                   # a VM error: never recorded
            if op == "pow":
                lhs = int(rng.integers(0, P, dtype=np.uint64)) if lhs > 3 else lhs      # the base is any field element
            if op in ("log_2_floor", "pop_count"):
                rhs = 0
            entries.append((op, lhs, rhs, int(rng.integers(1, 1000))))
    entries = list({(o, l, r): (o, l, r, m) for o, l, r, m in entries}.values())
    want_rows = []
    for op, lhs, rhs, mult in entries:
        for r in tg._u32_section(op, lhs, rhs, mult):
            want_rows.append([r["flag"], r["bits"], tg.F.inv((r["bits"] - 33) % P), tg.OPCODES[op], r["lhs"], tg.inv_or_zero(r["lhs"]), r["rhs"],
                              tg.inv_or_zero(r["rhs"]), r["result"], r["mult"]])
    total = len(want_rows)
    n = 1 << max(8, (total - 1).bit_length())
    words = [tg.OP_HALT]
    arrays = tg.aet_arrays(words, tg.execute(words))
    arrays["u32_entries"] = np.array([[tg.OPCODES[o], l, r, m] for o, l, r, m in entries], dtype=np.uint64)
    got, lengths = host_main_table(host, arrays, n)
    assert lengths[8] == total
    want = np.array(want_rows, dtype=np.uint64).T                      # [10][total]
    assert np.array_equal(got[139:149, :total], want)
    last = want_rows[-1]                                               # padding (u32.rs:126-154)
    pad = got[139:149, total:]
    assert (pad[0] == 0).all() and (pad[1] == 0).all() and (pad[2] == tg.F.inv((-33) % P)).all() and (pad[3] == last[3]).all()
    assert (pad[4] == last[4]).all() and (pad[5] == last[5]).all() and (pad[6] == 0).all() and (pad[7] == 0).all() and (pad[9] == 0).all()
    assert (pad[8] == (2 if last[3] == tg.OPCODES["lt"] else last[8])).all()

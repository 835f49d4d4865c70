"""The oracle's VM restatement (oracle/tracegen.execute: all 46 instructions) and the fill of ALL nine tables from an
execution — u32 sections, RAM table with its Bezout coefficients, hash table in all three modes, op-stack underflow, jump
stack, cascade / lookup multiplicities — checked against the one authority available without the Rust toolchain: the AIR.
Two programs that together execute every instruction yield tables on which all 604 constraints vanish, whose
cross-table arguments close, and whose proof verifies including the out-of-domain AIR identity."""
import numpy as np
import pytest

from conftest import rand_bfes
from oracle import corc, field as F, stark as S, tracegen as tg
from test_fibonacci_program import tables

P = F.P
KITCHEN_SINK = """
    push 1234567 push 89 lt pop 1
    push 61680 push 65280 and pop 1
    push 12345 push 54321 xor pop 1
    push 1000 log_2_floor pop 1
    push 10 push 3 pow pop 1
    push 7 push 100 div_mod pop 2
    push 255 pop_count pop 1
    push 18446744069414584320 split pop 2
    push 4294967299 split pop 2
    push 0 push 0 lt pop 1
    push 5 invert addi 7 pop 1
    push 1 push 2 push 3 push 4 push 5 push 6 xx_add
    push 7 push 8 push 9 xx_mul x_invert push 11 xb_mul pop 3
    divine 2 add
    push 2 swap 1 push 0 push 0 push 0 push 0 push 0
    call count_up
    pop 5 pop 2
    push 11 push 22 push 33 push 100 write_mem 3 pop 1
    push 102 read_mem 2 pop 3
    push 555 read_mem 1 pop 2
    push 44 push 101 write_mem 1 read_mem 1 pop 2
    push 1 push 2 push 3 push 4 push 5 push 6 push 7 push 8 push 9 push 10 hash
    dup 4 dup 4 dup 4 dup 4 dup 4 assert_vector
    sponge_init
    push 0 push 0 push 0 push 0 push 0 sponge_absorb
    sponge_squeeze pop 5 pop 5
    read_io 1 write_io 1
    nop halt
  count_up:
    pick 5 addi 1 place 5 recurse_or_return
"""
SINK_INPUT, SINK_SECRET = [42], [0, 0]

# Horner evaluation from memory, absorbing from memory, a two-level Merkle walk (one sibling divined, one from memory)
MEMORY_AND_MERKLE = """
    push 0 push 0 push 0 push 0 push 202 push 0 push 0 push 4 push 3 push 2
    b_horner_step b_horner_step b_horner_step
    pop 5 pop 2 write_io 3
    push 0 push 0 push 0 push 0 push 305 push 0 push 0 push 4 push 3 push 2
    x_horner_step x_horner_step
    pop 5 pop 2 write_io 3
    sponge_init push 400 sponge_absorb_mem pop 1
    sponge_squeeze write_io 5 pop 5
    push 0 push 0 push 0 push 0 push 500 push 0 push 6
    push 15 push 14 push 13 push 12 push 11
    merkle_step merkle_step_mem
    write_io 5 pop 5 pop 2
    halt
"""
MEM_RAM = {200: 7, 201: 8, 202: 9, 300: 1, 301: 2, 302: 3, 303: 4, 304: 5, 305: 6,
           **{400 + k: 100 + k for k in range(10)}, **{500 + k: 900 + k for k in range(5)}}
MEM_SIBLING = [21, 22, 23, 24, 25]


def _run(src, inp=(), sec=()):
    ex = tg.execute(tg.assemble(src), inp, sec)
    return ex


def test_instruction_semantics():
    assert _run("push 89 push 1234567 lt write_io 1 halt").output == [0]          # st0 < st1 ?
    assert _run("push 1234567 push 89 lt write_io 1 halt").output == [1]
    assert _run("push 61680 push 65280 and push 12345 push 54321 xor write_io 2 halt").output == [12345 ^ 54321, 61680 & 65280]
    assert _run("push 10 push 3 pow push 7 push 100 div_mod write_io 3 halt").output == [100 % 7, 100 // 7, 3 ** 10]
    assert _run("push 1000 log_2_floor push 255 pop_count write_io 2 halt").output == [8, 9]
    assert _run("push 18446744069414584320 split write_io 2 halt").output == [0, (1 << 32) - 1]      # lo, hi of p - 1
    assert _run("push 5 invert push 5 mul write_io 1 halt").output == [1]
    x, y = (4, 5, 6), (7, 8, 9)
    got = _run("push 6 push 5 push 4 push 9 push 8 push 7 xx_mul write_io 3 halt").output
    assert tuple(got) == F.xmul(x, y)
    assert _run("push 33 push 22 push 11 push 100 write_mem 3 pop 1 push 102 read_mem 3 pop 1 write_io 3 halt").output == [11, 22, 33]
    assert _run("push 77 read_mem 1 pop 1 write_io 1 halt").output == [0]                            # untouched RAM reads 0
    h = _run("push 10 push 9 push 8 push 7 push 6 push 5 push 4 push 3 push 2 push 1 hash write_io 5 halt").output
    from oracle import tip5
    assert h == [int(v) for v in tip5.hash_10(list(range(1, 11)))]                                   # st0 is the first input word
    assert _run("divine 2 add write_io 1 halt", sec=[3, 4]).output == [7]
    assert _run(KITCHEN_SINK, SINK_INPUT, SINK_SECRET).output == [42]


@pytest.mark.parametrize("src,exc", [("push 18446744069414584320 push 1 lt halt", "u32"), ("push 0 push 5 div_mod halt", "zero"),
                                     ("push 0 invert halt", "zero"), ("push 0 assert halt", "assert"),
                                     ("sponge_squeeze halt", "sponge"), ("pop 1 halt", "shallow")])
def test_instruction_errors(src, exc):
    with pytest.raises(ValueError, match=exc):
        _run(src)


def test_memory_and_merkle_program():
    from oracle import tip5
    words = tg.assemble(MEMORY_AND_MERKLE)
    ex = tg.execute(words, [], [], MEM_RAM, [MEM_SIBLING])
    x = (2, 3, 4)
    assert tuple(ex.output[:3]) == F.xadd(F.xmul(F.xadd(F.xmul((9, 0, 0), x), (8, 0, 0)), x), (7, 0, 0))      # 7 + 8x + 9x^2
    assert tuple(ex.output[3:6]) == F.xadd(F.xmul((4, 5, 6), x), (1, 2, 3))
    d1 = [int(v) for v in tip5.hash_pair([11, 12, 13, 14, 15], MEM_SIBLING)]        # node 6 is a left child
    assert ex.output[-5:] == [int(v) for v in tip5.hash_pair([900, 901, 902, 903, 904], d1)]   # node 3 is a right child
    with pytest.raises(IndexError):
        tg.execute(words, [], [], MEM_RAM, [])
    n = tg.padded_height(words, [], [], MEM_RAM, [MEM_SIBLING])
    T, digest, out = tg.main_table(words, [], n, [], MEM_RAM, [MEM_SIBLING])
    rng = np.random.default_rng(6)
    sampled = [tuple(int(v) for v in rng.integers(0, P, 3, dtype=np.uint64)) for _ in range(59)]
    ch = S.derive_challenges(sampled, S.Claim(digest, [], list(out)))
    B = corc.aux_extend(np.array(T, dtype=np.uint64), ch)
    assert tg.failing_constraints(T, [[tuple(int(v) for v in B[q][i]) for i in range(n)] for q in range(91)], ch) == []
    executed = {tg._NAME[r["ci"]] for r in ex.rows} | {tg._NAME[r["ci"]] for r in _run(KITCHEN_SINK, SINK_INPUT, SINK_SECRET).rows}
    assert executed | {"skiz", "return", "recurse", "assert", "eq", "mul"} == set(tg.OPCODES)   # the rest: test_fibonacci_program


def test_u32_sections_and_bezout_coefficients():
    sec = tg._u32_section("lt", 5, 9, 3)                          # u32.rs:193-290
    assert [r["bits"] for r in sec] == [0, 1, 2, 3, 4] and sec[0]["result"] == 1 and sec[0]["mult"] == 3 and sec[-1]["result"] == 2
    assert tg._u32_section("pow", 3, 10, 1)[0]["result"] == 3 ** 10
    assert tg._u32_section("log_2_floor", 1000, 0, 1)[0]["result"] == 9
    assert tg._u32_section("lt", 0, 0, 1)[0]["result"] == 0
    roots = [3, 5, 11, 1 << 40]
    a, b = tg.bezout_coefficients(roots)                          # a * rp + b * rp' == 1   (ram.rs:162-214)
    rp = [1]
    for r in roots:
        rp = tg._poly_mul(rp, [(-r) % P, 1])
    fd = [(k * rp[k]) % P for k in range(1, len(rp))]
    lhs = [0] * 8
    for k, c in enumerate(tg._poly_mul(a, rp)):
        lhs[k] = (lhs[k] + c) % P
    for k, c in enumerate(tg._poly_mul(b, fd)):
        lhs[k] = (lhs[k] + c) % P
    assert lhs == [1] + [0] * 7


def _sink_challenges(digest, out, seed=5):
    rng = np.random.default_rng(seed)
    sampled = [tuple(int(v) for v in rng.integers(0, P, 3, dtype=np.uint64)) for _ in range(59)]
    return S.derive_challenges(sampled, S.Claim(digest, list(SINK_INPUT), list(out)))


def test_kitchen_sink_satisfies_every_constraint():
    T, digest, out, ph, main = tables(KITCHEN_SINK, SINK_INPUT, None, SINK_SECRET)
    words = tg.assemble(KITCHEN_SINK)
    heights = tg.table_heights(words, tg.execute(words, SINK_INPUT, SINK_SECRET))
    assert ph == 2048 and heights["u32"] > 100 and heights["ram"] == 8 and heights["hash"] > 6 * 22
    ch = _sink_challenges(digest, out)
    B = corc.aux_extend(main, ch)
    n = main.shape[1]
    A = [[tuple(int(v) for v in B[q][i]) for i in range(n)] for q in range(91)]
    assert tg.failing_constraints(T, A, ch) == []
    from airgen.columns import MAIN
    for tab, col, row in (("u32", "Result", 3), ("ram", "RamValue", 2), ("hash", "State5", 140), ("processor", "HV0", 40),
                          ("ram", "BezoutCoefficientPolynomialCoefficient0", 1), ("op_stack", "FirstUnderflowElement", 7)):
        T2 = T.copy()
        idx = getattr(MAIN[tab], col)
        T2[idx, row] = (int(T2[idx, row]) + 1) % P
        assert tg.failing_constraints(T2, A, ch) != [], (tab, col)


def sink_instance(security, ldt, seed=31):
    st = S.Stark(security, 2, ldt)
    ph = tables(KITCHEN_SINK, SINK_INPUT, None, SINK_SECRET)[3]
    d = st.derive(ph)
    T, digest, out, _, main = tables(KITCHEN_SINK, SINK_INPUT, d["trace_len"], SINK_SECRET)
    n, h = d["trace_len"], d["num_trace_randomizers"]
    rng = np.random.default_rng(seed)
    mrand, arand, rcol = rand_bfes(rng, (379, h)), rand_bfes(rng, (91, h, 3)), rand_bfes(rng, (n, 3))
    qrand = rand_bfes(rng, (d["num_quotient_randomizer_coefficients"], 3))

    def cpu_extend(ch):
        return corc.aux_extend(main, np.asarray(ch, dtype=np.uint64).reshape(63, 3), rcol), arand
    return st, S.Claim(digest, list(SINK_INPUT), list(out)), main, mrand, cpu_extend, qrand, ph, rcol, arand


def test_oracle_proof_of_kitchen_sink_verifies_with_air_check():
    st, claim, main, mrand, cpu_extend, qrand, ph, _, _ = sink_instance(4, "fri")
    proof, _ = S.prove(st, claim, main, mrand, cpu_extend, qrand, padded_height=ph)
    assert S.verify(st, claim, proof, check_air=True)


@pytest.mark.gpu
def test_gpu_kitchen_sink_device_extension_and_proof(backend):
    import tvm_b200
    st, claim, main, mrand, cpu_extend, qrand, ph, rcol, arand = sink_instance(4, "fri")
    ch = np.array(_sink_challenges(claim.program_digest, claim.output), dtype=np.uint64)
    assert np.array_equal(backend.aux_extend(main, ch, rcol), corc.aux_extend(main, ch, rcol))   # real RAM / u32 / hash rows

    def device_extend(c):
        return backend.aux_extend(main, np.asarray(c, dtype=np.uint64).reshape(63, 3), rcol), arand
    got = backend.prove((claim.program_digest, claim.input, claim.output), main, mrand, device_extend, qrand,
                        security_level=4, log2_expansion=2, padded_height=ph, ldt_choice=tvm_b200.LDT_FRI)
    got = [int(v) for v in got]
    assert S.verify(st, claim, got, check_air=True)
    want, _ = S.prove(st, claim, main, mrand, cpu_extend, qrand, padded_height=ph)
    assert got == want


# ---- the reference's workload generator for an exact padded height: ProgramToBench::spin (triton-dev-util/src/lib.rs:49-75) ----
SPIN = "read_io 1 addi -3 push 2 pow place 5 call spin halt spin: pick 5 addi -1 place 5 recurse_or_return"


def spin_instance(log2_padded_height, security, ldt, seed=41):
    return program_instance(SPIN, [log2_padded_height], security, ldt, seed)


def program_instance(program, inp, security, ldt, seed=41, ram=None):
    words = tg.assemble(program)
    inp = list(inp)
    ph = tg.padded_height(words, inp, [], ram) if ram else tg.padded_height(words, inp)
    st = S.Stark(security, 2, ldt)
    d = st.derive(ph)
    n, h = d["trace_len"], d["num_trace_randomizers"]
    T, digest, out = tg.main_table(words, inp, n, [], ram) if ram else tg.main_table(words, inp, n)
    main = np.array(T, dtype=np.uint64)
    rng = np.random.default_rng(seed)
    mrand, arand, rcol = rand_bfes(rng, (379, h)), rand_bfes(rng, (91, h, 3)), rand_bfes(rng, (n, 3))
    qrand = rand_bfes(rng, (d["num_quotient_randomizer_coefficients"], 3))
    return dict(stark=st, claim=S.Claim(digest, inp, list(out)), main=main, T=T, main_rand=mrand, aux_rand=arand,
                randomizer_column=rcol, quot_rand=qrand, padded_height=ph, derived=d)


def test_spin_reaches_the_requested_padded_height_and_satisfies_the_air():
    import tvm_b200
    for k in (8, 9, 11):
        assert tg.padded_height(tg.assemble(SPIN), [k]) == 1 << k
    inst = spin_instance(9, 8, "fri")
    main, claim = inst["main"], inst["claim"]
    ch = _sink_challenges(claim.program_digest, claim.output)            # any challenges; the claim terms come from `claim`
    ch = S.derive_challenges(ch[:59], claim)
    B = corc.aux_extend(main, ch)
    n = main.shape[1]
    assert tg.failing_constraints(inst["T"], [[tuple(int(v) for v in B[q][i]) for i in range(n)] for q in range(91)], ch) == []
    rcol, arand = inst["randomizer_column"], inst["aux_rand"]
    proof, _ = S.prove(inst["stark"], claim, main, inst["main_rand"],
                       lambda c: (corc.aux_extend(main, np.asarray(c, dtype=np.uint64).reshape(63, 3), rcol), arand), inst["quot_rand"],
                       padded_height=inst["padded_height"])
    assert tvm_b200.verify((claim.program_digest, claim.input, claim.output), proof, 8, 2, ldt_choice=tvm_b200.LDT_FRI) == (True, "")


# A verifier-shaped program as the stand-in for BASELINE's "recursive-verifier program" (its TASM lives in the external tasm-lib,
# SURVEY 8(d).5): per iteration eight 10-word absorptions from memory (the hash table grows by 48 rows), u32 comparisons and
# bit operations, a division, X-field multiplication / inversion, RAM writes and reads - the instruction mix of a STARK
# verifier (Merkle/sponge hashing, index arithmetic, extension-field arithmetic), sized by its public input.
# (sponge_absorb_mem overwrites st1..st4 with the first absorbed words, vm.rs:699-729: four scratch words sit under the pointer)
VERIFIER_LIKE = """
    read_io 1 sponge_init call body sponge_squeeze write_io 5 pop 5 pop 1 halt
  body:
    dup 0 push 0 eq skiz return
    push 0 push 0 push 0 push 0 push 400 sponge_absorb_mem sponge_absorb_mem sponge_absorb_mem sponge_absorb_mem pop 5
    push 0 push 0 push 0 push 0 push 440 sponge_absorb_mem sponge_absorb_mem sponge_absorb_mem sponge_absorb_mem pop 5
    push 1234567 push 89 lt pop 1
    push 61680 push 65280 and push 12345 xor pop 1
    push 7 push 100 div_mod pop 2
    push 1000 log_2_floor pop_count pop 1
    push 1 push 2 push 3 push 4 push 5 push 6 xx_mul x_invert push 11 xb_mul
    push 100 write_mem 3 pop 1
    push 102 read_mem 3 pop 4
    addi -1 recurse
"""
VERIFIER_LIKE_RAM = {400 + k: 1000 + 7 * k for k in range(80)}


def _workload(name):
    """(program, public input, initial RAM) of a named workload: spin_K, fib_N (benches/prove_fib.rs:8-28), verifier_N"""
    from test_fibonacci_program import FIBONACCI
    kind, arg = name.split("_")
    if kind == "verifier":
        return VERIFIER_LIKE, [int(arg)], VERIFIER_LIKE_RAM
    return (SPIN, [int(arg)], None) if kind == "spin" else (FIBONACCI, [int(arg)], None)


def verifier_like_instance(iterations, security, ldt, seed=41):
    return program_instance(VERIFIER_LIKE, [iterations], security, ldt, seed, ram=VERIFIER_LIKE_RAM)


def test_verifier_like_program_runs_and_satisfies_the_air():
    words = tg.assemble(VERIFIER_LIKE)
    ex = tg.execute(words, [3], [], VERIFIER_LIKE_RAM)
    assert len(ex.output) == 5
    used = {tg._NAME[r["ci"]] for r in ex.rows}
    assert {"sponge_absorb_mem", "lt", "and", "xor", "div_mod", "log_2_floor", "pop_count", "xx_mul", "x_invert", "xb_mul", "write_mem",
            "read_mem", "recurse"} <= used
    inst = verifier_like_instance(3, 8, "fri")
    n = inst["main"].shape[1]
    rng = np.random.default_rng(3)
    sampled = [tuple(int(v) for v in rng.integers(0, P, 3, dtype=np.uint64)) for _ in range(59)]
    ch = S.derive_challenges(sampled, inst["claim"])
    B = corc.aux_extend(inst["main"], ch)
    assert tg.failing_constraints(inst["T"], [[tuple(int(v) for v in B[q][i]) for i in range(n)] for q in range(91)], ch) == []
    # sized for BASELINE heights by its input: 700 iterations pad to 2^16 (3 000 to 2^18, 11 500 to 2^20)
    assert tg.padded_height(words, [700], [], VERIFIER_LIKE_RAM) == 1 << 16


@pytest.mark.gpu
@pytest.mark.parametrize("workload,ldt", [("fib_100", None), ("spin_13", "fri"), ("spin_14", "stir"), ("spin_16", None),
                                          ("verifier_700", None),      # verifier-shaped stand-in for the recursive verifier, 2^16
                                          ("spin_18", None), ("spin_20", None)])   # None: Stark::default() — FRI below 2^16, STIR from
                                                                                 # there on; spin_20 is the BASELINE headline size
def test_gpu_proves_benchmark_workloads_from_the_149_table_columns(backend, workload, ldt):
    """The whole device-side pipeline on the reference's benchmark workloads (prove_fib, ProgramToBench::spin) at
    Stark::default() security, checked by the verifier INCLUDING the AIR — and, where tests/golden/spin_digests.json holds the
    oracle's digest of the same instance, word for word: degree-lowering main columns, auxiliary table, proof: all from the GPU."""
    import tvm_b200
    program, inp, ram = _workload(workload)
    inst = program_instance(program, inp, 160, ldt, ram=ram)
    log2_padded_height = inst["padded_height"].bit_length() - 1
    claim, want_main = inst["claim"], inst["main"]
    main = want_main.copy()
    main[149:] = 0
    backend.fill_derived_main_columns(main)
    assert np.array_equal(main, want_main)
    rcol, arand = inst["randomizer_column"], inst["aux_rand"]
    choice = {"stir": tvm_b200.LDT_STIR, "fri": tvm_b200.LDT_FRI, None: tvm_b200.LDT_AUTO}[ldt]
    # Stark::ldt (stark.rs:1942-1958): the forced choice, else FRI below padded height 2^16 and STIR from there on
    assert inst["derived"]["ldt"] == (ldt or ("fri" if log2_padded_height < 16 else "stir"))
    got = backend.prove((claim.program_digest, claim.input, claim.output), main, inst["main_rand"],
                        lambda ch: (backend.aux_extend(main, ch, rcol), arand), inst["quot_rand"], security_level=160,
                        log2_expansion=2, padded_height=inst["padded_height"], ldt_choice=choice)
    assert tvm_b200.proof_padded_height(got) == 1 << log2_padded_height
    # the same instance through tvm_prove_tables: derived columns and MasterMainTable::extend INSIDE the prove, on the resident
    # main trace (no callback, columns 149.. never uploaded) - identical proof words
    blank = want_main.copy()
    blank[149:] = 0
    got_tables = backend.prove_tables((claim.program_digest, claim.input, claim.output), blank, inst["main_rand"], arand, rcol,
                                      inst["quot_rand"], security_level=160, log2_expansion=2, padded_height=inst["padded_height"],
                                      ldt_choice=choice)
    assert np.array_equal(got_tables, got)
    # where the oracle's proof of this very instance has been computed once (tests/golden/make_spin_golden.py, minutes of CPU
    # time at 2^16), the GPU's proof must hash to the same digest: word-for-word parity at a BASELINE configuration
    import json, os
    from oracle import reference_prover as RP
    fixtures = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "spin_digests.json")))
    fixture = fixtures.get(workload)
    if fixture and fixture["ldt"] == inst["derived"]["ldt"]:
        assert len(got) == fixture["proof_words"] and RP.proof_digest([int(v) for v in got]) == fixture["tip5_digest"]
    assert tvm_b200.verify((claim.program_digest, claim.input, claim.output), got, 160, 2, ldt_choice=choice) == (True, "")
    assert tvm_b200.verify((claim.program_digest, claim.input, [1]), got, 160, 2, ldt_choice=choice)[0] is False


# ---- more of the reference's example programs (triton-dev-util/src/example_programs.rs:40-92; token streams) -----------------
GCD = """read_io 2 dup 1 dup 1 lt skiz swap 1
  loop_cond: dup 1 push 0 eq skiz call terminate dup 1 dup 1 div_mod swap 2 pop 2 swap 1 call loop_cond
  terminate: write_io 1 halt"""
MANY_U32 = """push 1311768464867721216 split push 13387 push 78810 lt push 5 push 7 pow push 69584 push 6796 xor
  push 64972 push 3915 and push 98668 push 15787 div_mod push 15787 push 98668 div_mod push 98141 push 7397 and
  push 67749 push 60797 lt push 49528 split push 53483 call lsb push 79655 call is_u32 push 60615 log_2_floor
  push 13 push 5 pow push 86323 push 37607 xor push 32374 push 20636 pow push 97416 log_2_floor
  push 14392 push 31589 div_mod halt
  lsb: push 2 swap 1 div_mod return
  is_u32: split pop 1 push 0 eq return"""


@pytest.mark.parametrize("program,inp,want", [(GCD, [42, 56], [14]), (GCD, [56, 42], [14]), (GCD, [17, 5], [1]), (MANY_U32, [], [])])
def test_example_programs_satisfy_the_air(program, inp, want):
    import math
    words = tg.assemble(program)
    ex = tg.execute(words, inp)
    assert ex.output == want and (not inp or want == [math.gcd(*inp)])
    n = tg.padded_height(words, inp)
    T, digest, out = tg.main_table(words, inp, n)
    rng = np.random.default_rng(8)
    sampled = [tuple(int(v) for v in rng.integers(0, P, 3, dtype=np.uint64)) for _ in range(59)]
    ch = S.derive_challenges(sampled, S.Claim(digest, list(inp), list(out)))
    B = corc.aux_extend(np.array(T, dtype=np.uint64), ch)
    assert tg.failing_constraints(T, [[tuple(int(v) for v in B[q][i]) for i in range(n)] for q in range(91)], ch) == []

"""The auxiliary-table extension generated FROM THE AIR (triton-vm_b200/airgen/extend_gen.py: symbolic differentiation
of every constraint that mentions an auxiliary column gives the per-row affine map of that column) against an independent
derivation: oracle/tracegen.extend_by_solving, which solves the raw, un-lowered, per-instruction-COMBINED constraints
numerically row by row.  Both must produce the same 91 columns — on real programs, where the result also satisfies all
604 constraints, and on synthetic tables that walk through all 47 instructions, every table's case distinctions and
the padding sections (there the AIR is not satisfied, but each auxiliary recurrence is still well defined)."""
import numpy as np
import pytest

from oracle import corc, field as F, stark as S, tracegen as tg
from test_fibonacci_program import BRANCHY, FIBONACCI, tables

P = F.P


def _challenges(seed):
    rng = np.random.default_rng(seed)
    return [tuple(int(v) for v in rng.integers(0, P, 3, dtype=np.uint64)) for _ in range(63)]


def semi_valid_table(n, seed, npad):
    """random main table whose selector columns are well formed: opcodes with their instruction bits and argument
    decomposition, table modes, padding flags"""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "triton-vm_b200"))
    from airgen.columns import MAIN
    from airgen.isa import ALL_INSTRUCTIONS, OPCODE
    rng = np.random.default_rng(seed)
    T = rng.integers(0, P, size=(379, n), dtype=np.uint64).astype(object)
    c = MAIN["processor"]
    ops = [OPCODE[i] for i in ALL_INSTRUCTIONS]
    for i in range(n):
        op = ops[i % len(ops)] if i < 2 * len(ops) else int(rng.choice(ops))
        T[c.CI, i] = op
        for b in range(7):
            T[c.IB0 + b, i] = (op >> b) & 1
        arg = int(rng.integers(1, 6))
        T[c.NIA, i] = arg
        for k in range(4):
            T[c.HV0 + k, i] = (arg >> k) & 1
        T[c.HV5, i] = int(rng.integers(0, 2))
        T[c.IsPadding, i] = 1 if i >= n - npad else 0
        T[c.CLK, i] = i
    h = MAIN["hash"]
    hops = [OPCODE[k] for k in ("hash", "sponge_init", "sponge_absorb", "sponge_squeeze")]
    for i in range(n):
        T[h.Mode, i], T[h.RoundNumber, i], T[h.CI, i] = int(rng.integers(0, 4)), int(rng.integers(0, 6)), int(rng.choice(hops))
    T[MAIN["ram"].InstructionType, :] = rng.integers(0, 3, n)
    T[MAIN["op_stack"].IB1ShrinkStack, :] = rng.integers(0, 3, n)
    T[MAIN["cascade"].IsPadding, :] = (np.arange(n) >= n - npad).astype(int)
    T[MAIN["lookup"].IsPadding, :] = (np.arange(n) >= n - npad).astype(int)
    u = MAIN["u32"]
    T[u.CopyFlag, :] = rng.integers(0, 2, n)
    T[u.CI, :] = rng.choice([OPCODE[k] for k in ("split", "lt", "and", "pow", "log_2_floor", "pop_count")], n)
    p = MAIN["program"]
    for i in range(n):
        T[p.IndexInChunk, i] = i % 10
        T[p.MaxMinusIndexInChunkInv, i] = tg.inv_or_zero(9 - i % 10)
        T[p.IsHashInputPadding, i] = 1 if i >= n - 2 * npad else 0
        T[p.IsTablePadding, i] = 1 if i >= n - npad else 0
    for name, col in (("ram", "RamPointer"), ("op_stack", "StackPointer"), ("jump_stack", "JSP")):
        cc = getattr(MAIN[name], col)                 # runs of equal pointers: the "same memory cell" branches
        for i in range(1, n):
            if rng.integers(0, 2):
                T[cc, i] = T[cc, i - 1]
    return T


@pytest.mark.parametrize("program,inp", [("halt", []), (FIBONACCI, [7]), (BRANCHY, [5, 7])])
def test_generated_extension_equals_air_solution_on_programs(program, inp):
    T, digest, out, _, main = tables(program, inp)
    rng = np.random.default_rng(5)
    sampled = [tuple(int(v) for v in rng.integers(0, P, 3, dtype=np.uint64)) for _ in range(59)]
    ch = S.derive_challenges(sampled, S.Claim(digest, list(inp), list(out)))
    A, unconstrained = tg.extend_by_solving(T, ch)
    assert unconstrained == []
    A = np.array(A, dtype=np.uint64)
    B = corc.aux_extend(main, ch, A[90])
    assert np.array_equal(A, B)
    n = main.shape[1]
    assert tg.failing_constraints(T, [[tuple(int(v) for v in B[q][i]) for i in range(n)] for q in range(91)], ch) == []


@pytest.mark.parametrize("seed,npad", [(1, 20), (2, 0)])
def test_generated_extension_equals_air_solution_on_every_instruction(seed, npad):
    n = 128
    T = semi_valid_table(n, seed, npad)
    ch = _challenges(50 + seed)
    A = np.array(tg.extend_by_solving(T, ch)[0], dtype=np.uint64)
    B = corc.aux_extend(np.array(T.tolist(), dtype=np.uint64), ch, A[90])
    assert np.array_equal(A, B)


def test_derived_columns_of_the_last_row_are_zero():
    # substitutions.rs:336-368: the dual-row rules run on rows 0..n-2 only
    T = semi_valid_table(64, 3, 5)
    B = corc.aux_extend(np.array(T.tolist(), dtype=np.uint64), _challenges(9))
    assert not B[49:90, -1].any() and B[49:90, :-1].any()
    assert not B[90].any()                          # no randomizer column supplied


@pytest.mark.parametrize("program,inp", [(FIBONACCI, [7]), (BRANCHY, [5, 7])])
def test_generated_derived_main_columns_equal_the_evaluated_substitutions(program, inp):
    # tracegen.fill_derived_main_columns evaluates the substitution circuits node by node; the generated straight-line
    # code (the text the device compiles) must reproduce all 230 columns from the 149 table columns
    T, _, _, _, main = tables(program, inp)                  # derived columns from the generated rules (the default)
    scrambled = main.copy()
    scrambled[149:] = 12345
    assert np.array_equal(corc.fill_derived_main(scrambled), main)
    T2 = T.copy()
    T2[149:, :] = 0
    tg.fill_derived_main_columns(T2)                         # node-by-node evaluation of the substitution circuits
    assert np.array_equal(np.array(T2.tolist(), dtype=np.uint64), main)
    assert main[149:169].any() and main[169:].any() and not main[169:, -1].any()      # last row of the tran section is 0

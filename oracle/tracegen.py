"""Trace generation for the smallest Triton VM program (`halt`): master main table (fill + pad + derived columns) and
master auxiliary table (`extend`), restated from the reference so that an AIR-SATISFYING instance exists in this
repository — with it the oracle verifier runs with `check_air=True` (the out-of-domain AIR / quotient identity,
stark.rs:1469-1540), which synthetic tables can never pass.

TEST INFRASTRUCTURE ONLY.  Follows, table by table:
  aet.rs:95-215 (program hashing trace, lookup multiplicities), vm.rs:246-268, 1113-1190 (initial state, processor row),
  table/master_table.rs:881-1004 (fill order, pad, derived columns), table/{program,processor,op_stack,ram,jump_stack,
  hash,cascade,lookup,u32}.rs (fill / pad / extend), triton-constraint-builder/src/substitutions.rs:128-330
  (derived-column fill).  Only what the one-instruction program exercises is restated: no op-stack underflow, RAM,
  u32, sponge or hash-instruction rows.
"""
import os
import sys

import numpy as np

from . import field as F, tip5
from .field import P, R

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "triton-vm_b200"))
from airgen.build import CATEGORIES, build_air            # noqa: E402
from airgen.columns import MAIN, AUX, CH                    # noqa: E402
from airgen.evaluate import evaluate_constraints            # noqa: E402

OP_HALT, OP_HASH, OP_SPLIT = 0, 18, 4
NUM_MAIN, NUM_AUX = 379, 91                                  # incl. derived columns; aux incl. the randomizer column
_AIR = None


def air():
    global _AIR
    if _AIR is None:
        _AIR = build_air()
    return _AIR


def inv_or_zero(x):
    x %= P
    return F.inv(x) if x else 0


# ---- Tip5 with trace (twenty-first Tip5::trace: state before round 0 .. after round 4) ----------------------
def tip5_trace(state):
    s = list(state)
    rows = [list(s)]
    for rnd in range(tip5.ROUNDS):
        for i in range(tip5.NUM_SPLIT_AND_LOOKUP):
            s[i] = tip5._split_and_lookup(s[i])
        for i in range(tip5.NUM_SPLIT_AND_LOOKUP, tip5.STATE):
            s[i] = pow(s[i], 7, P)
        t = [sum(tip5.MDS_FIRST_COLUMN[(r - c) % 16] * s[c] for c in range(16)) % P for r in range(16)]
        s = [(t[i] + tip5.ROUND_CONSTANTS[16 * rnd + i]) % P for i in range(16)]
        rows.append(list(s))
    return rows


def limbs16(x):                      # hash.rs:30-33: 16-bit limbs of the Montgomery representation
    m = x * R % P
    return [(m >> s) & 0xFFFF for s in (0, 16, 32, 48)]


def lookup8(v): return tip5.LOOKUP_TABLE[v]
def lookup16(v): return (lookup8(v >> 8) << 8) + lookup8(v & 0xFF)       # cascade.rs:29-35


# ---- main table -------------------------------------------------------------------------------------
def halt_main_table(n):
    """[379][n] canonical uint64: MasterMainTable::new + pad for the program `halt` (master_table.rs:881-1004)."""
    assert n >= 256 and n & (n - 1) == 0
    T = np.zeros((NUM_MAIN, n), dtype=object)
    program = [OP_HALT]

    # -- AET: program hashing (aet.rs:150-190) and the cascade / lookup multiplicities (305-344)
    padded_len = -(-(len(program) + 1) // 10) * 10
    padded_program = (program + [1] + [0] * 10)[:padded_len]
    cascade_mult, lookup_mult = {}, [0] * 256
    hash_rows = []
    sponge = [0] * 16
    for c0 in range(0, padded_len, 10):
        sponge[:10] = padded_program[c0:c0 + 10]
        trace = tip5_trace(sponge)
        for row in trace[:-1]:
            for e in row[:4]:
                for limb in limbs16(e):
                    if limb in cascade_mult:
                        cascade_mult[limb] += 1
                    else:
                        cascade_mult[limb] = 1
                        lookup_mult[limb & 0xFF] += 1
                        lookup_mult[limb >> 8] += 1
        for rnd, row in enumerate(trace):
            hash_rows.append((rnd, row))
        sponge = trace[-1]
    program_digest = sponge[:5]

    # -- program table (program.rs:33-113)
    c = MAIN["program"]
    for i in range(n):
        T[c.Address, i] = i
        T[c.IndexInChunk, i] = i % 10
        T[c.MaxMinusIndexInChunkInv, i] = inv_or_zero(9 - i % 10)
        if i < padded_len:
            T[c.Instruction, i] = padded_program[i]
            T[c.LookupMultiplicity, i] = 1 if i < len(program) else 0     # `halt` is executed once
            T[c.IsHashInputPadding, i] = 0 if i < len(program) else 1
        else:
            T[c.IsHashInputPadding, i] = 1
            T[c.IsTablePadding, i] = 1

    # -- processor table (vm.rs:1113-1190, processor.rs:45-95): one executed row, then padding
    c = MAIN["processor"]
    row = {c.CLK: 0, c.IP: 0, c.CI: OP_HALT, c.NIA: 1, c.OpStackPointer: 16}
    for b in range(7):
        row[c.IB0 + b] = (OP_HALT >> b) & 1
    # OpStack::new (op_stack.rs:58-68): the reversed digest occupies the 5 DEEPEST of the 16 stack registers, i.e.
    # st11..st15 = digest[0..5]; st0..st10 = 0
    for i in range(16):
        row[c.ST0 + i] = program_digest[i - 11] if i >= 11 else 0
    for k, v in row.items():
        T[k, 0] = v
    for i in range(1, n):
        for k, v in row.items():
            T[k, i] = v
        T[c.IsPadding, i] = 1
        T[c.CLK, i] = i
    T[c.ClockJumpDifferenceLookupMultiplicity, 1] = (T[c.ClockJumpDifferenceLookupMultiplicity, 1] + (n - 1)) % P

    # -- op stack table: empty -> padding rows only (op_stack.rs:197-211)
    c = MAIN["op_stack"]
    T[c.IB1ShrinkStack, :] = 2
    T[c.StackPointer, :] = 16

    # -- RAM table: empty (ram.rs:89-103)
    c = MAIN["ram"]
    T[c.InstructionType, :] = 2
    T[c.BezoutCoefficientPolynomialCoefficient1, :] = 1

    # -- jump stack table (jump_stack.rs:90-205)
    c = MAIN["jump_stack"]
    for i in range(n):
        T[c.CLK, i] = i
        T[c.CI, i] = OP_HALT

    # -- hash table (hash.rs:36-302)
    c = MAIN["hash"]
    names = c.names

    def col(name): return c.start + names.index(name)
    parts = ("Lowest", "MidLow", "MidHigh", "Highest")
    for i, (rnd, st) in enumerate(hash_rows):
        T[c.Mode, i] = 1
        T[c.CI, i] = OP_HASH
        T[c.RoundNumber, i] = rnd
        for e in range(4):
            lb = limbs16(st[e])
            for k, part in enumerate(parts):
                T[col(f"State{e}{part}LkIn"), i] = lb[k]
                T[col(f"State{e}{part}LkOut"), i] = lookup16(lb[k])
            T[col(f"State{e}Inv"), i] = inv_or_zero((1 << 32) - 1 - ((lb[3] << 16) + lb[2]))
        for e in range(4, 16):
            T[col(f"State{e}"), i] = st[e]
        for k in range(16):
            T[col(f"Constant{k}"), i] = tip5.ROUND_CONSTANTS[16 * rnd + k] if rnd < 5 else 0
    zero_inv = inv_or_zero((1 << 32) - 1)
    for i in range(len(hash_rows), n):
        for e in range(4):
            T[col(f"State{e}Inv"), i] = zero_inv
        for k in range(16):
            T[col(f"Constant{k}"), i] = tip5.ROUND_CONSTANTS[k]
        T[c.Mode, i] = 0
        T[c.CI, i] = OP_HASH

    # -- cascade table (cascade.rs:41-66): insertion order of the multiplicity map
    c = MAIN["cascade"]
    for i, (limb, mult) in enumerate(cascade_mult.items()):
        T[c.LookInLo, i] = limb & 0xFF
        T[c.LookInHi, i] = limb >> 8
        T[c.LookOutLo, i] = lookup8(limb & 0xFF)
        T[c.LookOutHi, i] = lookup8(limb >> 8)
        T[c.LookupMultiplicity, i] = mult
    for i in range(len(cascade_mult), n):
        T[c.IsPadding, i] = 1

    # -- lookup table (lookup.rs:84-116)
    c = MAIN["lookup"]
    for i in range(256):
        T[c.LookIn, i] = i
        T[c.LookOut, i] = lookup8(i)
        T[c.LookupMultiplicity, i] = lookup_mult[i]
    for i in range(256, n):
        T[c.IsPadding, i] = 1

    # -- u32 table: empty (u32.rs:126-154)
    c = MAIN["u32"]
    T[c.CI, :] = OP_SPLIT
    T[c.BitsMinus33Inv, :] = F.inv((-33) % P)

    fill_derived_main_columns(T)
    return T, program_digest


def _derive(constraints, start, T_cur_row, T_next_row, aux_cur, aux_next, challenges, is_main):
    """value of derived column k = the substituted expression = -(constraint with the new column set to 0)"""
    out = []
    cur = list(T_cur_row) if is_main else list(aux_cur)
    for k, cnode in enumerate(constraints):
        if is_main:
            cur_main, cur_aux = cur + [0] * (NUM_MAIN - len(cur)), aux_cur
        else:
            cur_main, cur_aux = T_cur_row, cur + [(0, 0, 0)] * (NUM_AUX - len(cur))
        v = evaluate_constraints([cnode], cur_main, cur_aux, T_next_row, aux_next, challenges)[0]
        val = F.xneg(v)
        if is_main:
            assert val[1] == 0 and val[2] == 0
            cur.append(val[0]); out.append(val[0])
        else:
            cur.append(val); out.append(val)
    return out


def fill_derived_main_columns(T):
    """DegreeLoweringTable::fill_derived_main_columns (substitutions.rs:128-161, 237-300): sections init | cons | tran
    | term; a transition-section column of the last row stays 0."""
    a = air()
    n = T.shape[1]
    zero_aux = [(0, 0, 0)] * NUM_AUX
    ch = [(0, 0, 0)] * 63
    for cat in CATEGORIES:
        rules = a.main_subst[cat]
        if not rules:
            continue
        start = a.subst_col_start[cat][0]
        dual = cat == "tran"
        for i in range(n - 1 if dual else n):
            cur = [int(T[q, i]) for q in range(start)]
            nxt = [int(T[q, i + 1]) for q in range(NUM_MAIN)] if dual else [0] * NUM_MAIN
            vals = _derive(rules, start, cur, nxt, zero_aux, zero_aux, ch, True)
            for k, v in enumerate(vals):
                T[start + k, i] = v


# ---- AIR check ---------------------------------------------------------------------------------------
def column_name(is_main, col):
    spec = MAIN if is_main else AUX
    base = 149 if is_main else 49
    if col >= base:
        return ("main" if is_main else "aux") + f".derived{col - base}"
    for tname, e in spec.items():
        if e.start <= col < e.start + e.COUNT:
            return f"{tname}.{e.names[col - e.start]}"
    return f"?{col}"


def failing_constraints(T, A, challenges, max_report=12):
    """Evaluates every constraint where it must vanish (master_table.rs:1194-1252 zerofiers): initial on row 0,
    consistency on all rows, transition on rows (i, i+1) for i < n-1, terminal on row n-1.
    T [379][n] ints, A [91][n] X-field tuples (column 90 = randomizer, ignored by the AIR).  Returns a list of
    (category, constraint index, row, referenced columns)."""
    from . import corc
    from airgen.circuit import reachable_postorder
    a = air()
    n = T.shape[1]
    rows_m = [[int(T[q, i]) for q in range(NUM_MAIN)] for i in range(n)]
    rows_a = [[tuple(int(v) for v in A[q][i]) for q in range(90)] for i in range(n)]
    bad = []

    def refs(cat, k):
        out = set()
        for nd in reachable_postorder([a.constraints[cat][k]]):
            if nd.kind == "I":
                r, is_main, col = nd.val
                out.add(("next." if r else "") + column_name(is_main, col))
        return sorted(out)

    def run(cat_idx, cat, pairs):
        for i, j in pairs:
            ev = corc.air_eval_category(cat_idx, rows_m[i], rows_a[i], rows_m[j], rows_a[j], challenges)
            for k in np.nonzero(ev.any(axis=1))[0]:
                bad.append((cat, int(k), i))
                if len(bad) >= max_report * 50:
                    return
    run(0, "init", [(0, 0)])
    run(1, "cons", [(i, i) for i in range(n)])
    run(2, "tran", [(i, i + 1) for i in range(n - 1)])
    run(3, "term", [(n - 1, n - 1)])
    seen, report = set(), []
    for cat, k, i in bad:
        if (cat, k) not in seen:
            seen.add((cat, k))
            report.append((cat, k, i, refs(cat, k)))
    return report


# ---- auxiliary table ---------------------------------------------------------------------------------
# The auxiliary columns are running products / evaluation arguments / logarithmic derivatives whose values the AIR
# itself pins down: the initial constraints fix row 0 and every transition constraint is affine in the next row's
# auxiliary value.  Instead of restating the nine `extend` functions, the table is obtained by SOLVING the (not yet
# degree-lowered) constraints column by column, row by row:  f(u) = f(0) + u (f(1) - f(0)) = 0.  The derived
# auxiliary columns then follow from the substitution rules like the main ones, and `failing_constraints` confirms
# the complete, lowered AIR on the result.
_RAW = None


def raw_air():
    """constraints before degree lowering: {category: (builder, [root nodes])}"""
    global _RAW
    if _RAW is None:
        from airgen.build import PROVIDERS, _FN
        from airgen.circuit import Builder
        _RAW = {}
        for cat in CATEGORIES:
            b = Builder(dual=(cat == "tran"))
            roots = []
            for prov in PROVIDERS:
                roots += [m.n for m in getattr(prov, _FN[cat])(b)]
            _RAW[cat] = (b, roots)
    return _RAW


class _Evaluator:
    """memoised post-order evaluation of one constraint"""

    def __init__(self, node):
        from airgen.circuit import reachable_postorder
        self.order = reachable_postorder([node])
        self.root = node
        self.inputs = {n.val for n in self.order if n.kind == "I"}

    def __call__(self, cur_main, cur_aux, next_main, next_aux, challenges):
        from airgen.circuit import xmul, xadd
        val = {}
        for n in self.order:
            k = n.kind
            if k == "B": v = (n.val, 0, 0)
            elif k == "X": v = n.val
            elif k == "C": v = challenges[n.val]
            elif k == "I":
                row, is_main, col = n.val
                if is_main:
                    v = ((next_main if row else cur_main)[col] % P, 0, 0)
                else:
                    v = (next_aux if row else cur_aux)[col]
            elif k == "+": v = xadd(val[id(n.lhs)], val[id(n.rhs)])
            else: v = xmul(val[id(n.lhs)], val[id(n.rhs)])
            val[id(n)] = v
        return val[id(self.root)]


def _solve_affine(ev, cur_main, cur_aux, next_main, next_aux, challenges, row_sel, q):
    """root of u -> constraint(u) where u is auxiliary column q of the current (row_sel=0) or next row; None if the
    constraint does not depend on u for these inputs"""
    target = next_aux if row_sel else cur_aux
    target[q] = (0, 0, 0)
    f0 = ev(cur_main, cur_aux, next_main, next_aux, challenges)
    target[q] = (1, 0, 0)
    f1 = ev(cur_main, cur_aux, next_main, next_aux, challenges)
    slope = F.xsub(f1, f0)
    if slope == (0, 0, 0):
        return None
    return F.xneg(F.xmul(f0, F.xinv(slope)))


def extend_by_solving(T, challenges, randomizer_seed=1):
    """-> A [91][n] of X-field tuples for main table T [379][n] and the 63 challenges"""
    a = air()
    raw = raw_air()
    n = T.shape[1]
    rows_m = [[int(T[q, i]) for q in range(NUM_MAIN)] for i in range(n)]
    A = [[(0, 0, 0)] * NUM_AUX for _ in range(n)]                    # row-major while solving
    init_evs = [_Evaluator(c) for c in raw["init"][1]]
    tran_evs = [_Evaluator(c) for c in raw["tran"][1]]
    base_aux = list(range(49))
    # row 0 from the initial constraints
    solved = set()
    progress = True
    while progress and len(solved) < 49:
        progress = False
        for q in base_aux:
            if q in solved: continue
            for ev in init_evs:
                aux_refs = {col for (r, is_main, col) in ev.inputs if not is_main}
                if q not in aux_refs or not (aux_refs - {q}) <= solved: continue
                u = _solve_affine(ev, rows_m[0], A[0], rows_m[0], A[0], challenges, 0, q)
                if u is not None:
                    A[0][q] = u; solved.add(q); progress = True
                    break
            else:
                A[0][q] = (0, 0, 0)
    unsolved0 = [q for q in base_aux if q not in solved]
    # rows 1.. from the transition constraints
    by_col = {q: [ev for ev in tran_evs if (1, False, q) in ev.inputs] for q in base_aux}
    unconstrained = set(unsolved0)
    for i in range(n - 1):
        solved = set()
        nxt = A[i + 1]
        progress = True
        while progress and len(solved) < 49:
            progress = False
            for q in base_aux:
                if q in solved: continue
                for ev in by_col[q]:
                    next_refs = {col for (r, is_main, col) in ev.inputs if not is_main and r == 1}
                    if not (next_refs - {q}) <= solved: continue
                    u = _solve_affine(ev, rows_m[i], A[i], rows_m[i + 1], nxt, challenges, 1, q)
                    if u is not None:
                        nxt[q] = u; solved.add(q); progress = True
                        break
        for q in base_aux:
            if q not in solved:                       # no transition constraint moves it on this row: carry over
                nxt[q] = A[i][q]
                unconstrained.add(q)
    # derived auxiliary columns (substitutions.rs:163-330)
    for cat in CATEGORIES:
        rules = a.aux_subst[cat]
        if not rules: continue
        start = a.subst_col_start[cat][1]
        dual = cat == "tran"
        evs = [_Evaluator(c) for c in rules]
        for i in range(n - 1 if dual else n):
            cur_aux = A[i]
            nm = rows_m[i + 1] if dual else rows_m[i]
            na = A[i + 1] if dual else A[i]
            for k, ev in enumerate(evs):
                cur_aux[start + k] = (0, 0, 0)
                v = ev(rows_m[i], cur_aux, nm, na, challenges)
                cur_aux[start + k] = F.xneg(v)
    rng = np.random.default_rng(randomizer_seed)
    for i in range(n):
        A[i][90] = tuple(int(v) for v in rng.integers(0, P, 3, dtype=np.uint64))
    return [[A[i][q] for i in range(n)] for q in range(NUM_AUX)], sorted(unconstrained)

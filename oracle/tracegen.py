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


# ---- a small VM (vm.rs:361-1000) ---------------------------------------------------------------------------
# Instruction subset: halt nop push pop dup swap add mul eq skiz call return recurse assert read_io write_io.
# No RAM, u32, hash or sponge instructions (their tables stay in the all-padding form).
from .isa_words import OPCODES, HAS_ARG, assemble   # noqa: E402

_NAME = {v: k for k, v in OPCODES.items()}
SUPPORTED = {"halt", "nop", "push", "pop", "dup", "swap", "add", "mul", "eq", "skiz", "call", "return", "recurse", "assert",
             "read_io", "write_io"}


def run(words, public_input=()):
    """-> (processor rows [dict], op-stack table entries [(clk, shrink, pointer, payload)], instruction multiplicities,
    public output, program digest)"""
    digest = [int(v) for v in tip5.hash_varlen(words)]
    stack = list(reversed(digest)) + [0] * 11                 # OpStack::new: vec index 0 = deepest element
    jump_stack, inp, out = [], list(public_input), []
    mult = [0] * len(words)
    rows, os_entries = [], []
    ip = clk = 0
    while True:
        if not 0 <= ip < len(words):
            raise ValueError("instruction pointer out of bounds")
        name = _NAME[words[ip]]
        if name not in SUPPORTED:
            raise ValueError(f"instruction {name} is not restated")
        arg = words[ip + 1] if name in HAS_ARG else None
        size = 2 if name in HAS_ARG else 1
        nxt_ip = ip + size
        nia = arg if arg is not None else (words[nxt_ip] if nxt_ip < len(words) else 1)      # vm.rs:1178-1189
        hv = [0] * 6                                                                        # vm.rs:270-345
        st = lambda i: stack[len(stack) - 1 - i]                                           # noqa: E731
        if name in ("pop", "dup", "swap", "read_io", "write_io"):
            hv[:4] = [(arg >> k) & 1 for k in range(4)]
        elif name == "skiz":
            hv[0] = inv_or_zero(st(0))
            hv[1:6] = [nia % 2, (nia >> 1) % 4, (nia >> 3) % 4, (nia >> 5) % 4, nia >> 7]
        elif name == "eq":
            hv[0] = inv_or_zero(st(1) - st(0))
        rows.append(dict(clk=clk, ip=ip, ci=words[ip], nia=nia, jsp=len(jump_stack),
                         jso=jump_stack[-1][0] if jump_stack else 0, jsd=jump_stack[-1][1] if jump_stack else 0,
                         st=[st(i) for i in range(16)], osp=len(stack), hv=hv))
        mult[ip] += 1
        io = []                                                   # underflow IO of this instruction (op_stack.rs:77-103)

        def push(e):
            stack.append(e % P)
            io.append(("w", stack[len(stack) - 1 - 16]))

        def pop():
            io.append(("r", stack[len(stack) - 1 - 16] if len(stack) > 16 else 0))
            if len(stack) <= 16:
                raise ValueError("op stack too shallow")
            return stack.pop()
        halting = False
        if name == "halt": halting = True; ip = nxt_ip
        elif name == "nop": ip = nxt_ip
        elif name == "push": push(arg); ip = nxt_ip
        elif name == "pop":
            for _ in range(arg): pop()
            ip = nxt_ip
        elif name == "dup": push(st(arg)); ip = nxt_ip
        elif name == "swap":
            i0, i1 = len(stack) - 1, len(stack) - 1 - arg
            stack[i0], stack[i1] = stack[i1], stack[i0]
            ip = nxt_ip
        elif name == "add": x = pop(); y = pop(); push(x + y); ip = nxt_ip
        elif name == "mul": x = pop(); y = pop(); push(x * y); ip = nxt_ip
        elif name == "eq": x = pop(); y = pop(); push(1 if x == y else 0); ip = nxt_ip
        elif name == "assert":
            if pop() != 1: raise ValueError("assertion failed")
            ip = nxt_ip
        elif name == "skiz":
            top = pop()
            if top == 0:
                nn = _NAME[words[nxt_ip]]
                ip = nxt_ip + (2 if nn in HAS_ARG else 1)
            else:
                ip = nxt_ip
        elif name == "call": jump_stack.append((ip + 2, arg)); ip = arg
        elif name == "return": ip = jump_stack.pop()[0]
        elif name == "recurse": ip = jump_stack[-1][1]
        elif name == "read_io":
            for _ in range(arg): push(inp.pop(0))
            ip = nxt_ip
        elif name == "write_io":
            for _ in range(arg): out.append(pop())
            ip = nxt_ip
        # canonicalise the underflow IO sequence and turn it into table entries (op_stack.rs:61-87, 234-255)
        changed = True
        while changed:
            changed = False
            for k in range(len(io) - 1):
                if io[k][0] != io[k + 1][0] and io[k][1] == io[k + 1][1]:
                    del io[k:k + 2]; changed = True
                    break
        assert len({t for t, _ in io}) <= 1
        if io:
            ptr = len(stack) - len(io) if io[0][0] == "w" else len(stack) + len(io)
            for t, payload in io:
                if t == "r":
                    ptr -= 1
                os_entries.append((clk, 1 if t == "r" else 0, ptr, payload))
                if t == "w":
                    ptr += 1
        clk += 1
        if halting:
            break
    return rows, os_entries, mult, out, digest


# ---- main table -------------------------------------------------------------------------------------
def main_table(words, public_input, n):
    """[379][n] canonical ints: MasterMainTable::new + pad (master_table.rs:881-1004) for a program of the supported
    instruction subset.  -> (table, program digest, public output)"""
    assert n >= 256 and n & (n - 1) == 0
    T = np.zeros((NUM_MAIN, n), dtype=object)
    program = list(words)
    rows, os_entries, mult, public_output, program_digest = run(program, public_input)
    plen = len(rows)
    assert plen <= n and len(os_entries) <= n

    # -- AET: program hashing (aet.rs:150-190) and the cascade / lookup multiplicities (305-344)
    padded_len = -(-(len(program) + 1) // 10) * 10
    padded_program = (program + [1] + [0] * 10)[:padded_len]
    assert padded_len <= n
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
        sponge = list(trace[-1])
    assert sponge[:5] == program_digest
    assert len(hash_rows) <= n and len(cascade_mult) <= n

    # -- program table (program.rs:33-113)
    c = MAIN["program"]
    for i in range(n):
        T[c.Address, i] = i
        T[c.IndexInChunk, i] = i % 10
        T[c.MaxMinusIndexInChunkInv, i] = inv_or_zero(9 - i % 10)
        if i < padded_len:
            T[c.Instruction, i] = padded_program[i]
            T[c.LookupMultiplicity, i] = mult[i] if i < len(program) else 0
            T[c.IsHashInputPadding, i] = 0 if i < len(program) else 1
        else:
            T[c.IsHashInputPadding, i] = 1
            T[c.IsTablePadding, i] = 1

    # -- op stack table (op_stack.rs:179-211): sorted by (stack pointer, clk); padding copies the last row
    c = MAIN["op_stack"]
    os_sorted = sorted(os_entries, key=lambda e: (e[2], e[0]))
    clk_jump_diffs = []
    for i, (clk, shrink, ptr, payload) in enumerate(os_sorted):
        T[c.CLK, i], T[c.IB1ShrinkStack, i], T[c.StackPointer, i], T[c.FirstUnderflowElement, i] = clk, shrink, ptr, payload
        if i and os_sorted[i - 1][2] == ptr:
            clk_jump_diffs.append(clk - os_sorted[i - 1][0])
    if os_sorted:
        last = len(os_sorted) - 1
        for i in range(len(os_sorted), n):
            T[c.CLK, i], T[c.StackPointer, i] = T[c.CLK, last], T[c.StackPointer, last]
            T[c.FirstUnderflowElement, i] = T[c.FirstUnderflowElement, last]
            T[c.IB1ShrinkStack, i] = 2
    else:
        T[c.IB1ShrinkStack, :] = 2
        T[c.StackPointer, :] = 16

    # -- RAM table: empty (ram.rs:89-103)
    c = MAIN["ram"]
    T[c.InstructionType, :] = 2
    T[c.BezoutCoefficientPolynomialCoefficient1, :] = 1

    # -- jump stack table (jump_stack.rs:90-205): grouped by jsp, execution order inside a group; the padding rows
    #    follow the row with the largest clock cycle, the rows after it move to the end
    c = MAIN["jump_stack"]
    groups = {}
    for r in rows:
        groups.setdefault(r["jsp"], []).append((r["clk"], r["ci"], r["jsp"], r["jso"], r["jsd"]))
    js = [e for jsp in sorted(groups) for e in groups[jsp]]
    for i in range(len(js) - 1):
        if js[i][2] == js[i + 1][2]:
            clk_jump_diffs.append(js[i + 1][0] - js[i][0])
    k_max = next(i for i, e in enumerate(js) if e[0] == plen - 1)
    padded_js = js[:k_max + 1] + [(clk, js[k_max][1], js[k_max][2], js[k_max][3], js[k_max][4]) for clk in range(plen, n)] + js[k_max + 1:]
    for i, (clk, ci, jsp, jso, jsd) in enumerate(padded_js):
        T[c.CLK, i], T[c.CI, i], T[c.JSP, i], T[c.JSO, i], T[c.JSD, i] = clk, ci, jsp, jso, jsd

    # -- processor table (vm.rs:1113-1190, processor.rs:45-95)
    c = MAIN["processor"]

    def put(i, r, padding):
        T[c.CLK, i], T[c.IP, i], T[c.CI, i], T[c.NIA, i] = r["clk"], r["ip"], r["ci"], r["nia"]
        for b in range(7):
            T[c.IB0 + b, i] = (r["ci"] >> b) & 1
        T[c.JSP, i], T[c.JSO, i], T[c.JSD, i] = r["jsp"], r["jso"], r["jsd"]
        for k in range(16):
            T[c.ST0 + k, i] = r["st"][k]
        T[c.OpStackPointer, i] = r["osp"]
        for k in range(6):
            T[c.HV0 + k, i] = r["hv"][k]
        T[c.IsPadding, i] = 1 if padding else 0
    for i, r in enumerate(rows):
        put(i, r, False)
    for d in clk_jump_diffs:
        T[c.ClockJumpDifferenceLookupMultiplicity, d] += 1
    for i in range(plen, n):
        put(i, dict(rows[-1], clk=i), True)
    if n > plen:
        T[c.ClockJumpDifferenceLookupMultiplicity, 1] = (T[c.ClockJumpDifferenceLookupMultiplicity, 1] + (n - plen)) % P

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
    for i, (limb, m) in enumerate(cascade_mult.items()):
        T[c.LookInLo, i] = limb & 0xFF
        T[c.LookInHi, i] = limb >> 8
        T[c.LookOutLo, i] = lookup8(limb & 0xFF)
        T[c.LookOutHi, i] = lookup8(limb >> 8)
        T[c.LookupMultiplicity, i] = m
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
    return T, program_digest, public_output


def padded_height(words, public_input=()):
    """AlgebraicExecutionTrace::padded_height (aet.rs:99-135) for the supported subset"""
    rows, os_entries, _, _, _ = run(list(words), public_input)
    padded_len = -(-(len(words) + 1) // 10) * 10
    hash_len = 6 * (padded_len // 10)
    cascade = set()
    sponge = [0] * 16
    pp = (list(words) + [1] + [0] * 10)[:padded_len]
    for c0 in range(0, padded_len, 10):
        sponge[:10] = pp[c0:c0 + 10]
        tr = tip5_trace(sponge)
        for row in tr[:-1]:
            for e in row[:4]:
                cascade.update(limbs16(e))
        sponge = list(tr[-1])
    h = max(padded_len, len(rows), len(os_entries), hash_len, len(cascade), 256)
    p2 = 1
    while p2 < h: p2 <<= 1
    return p2


def halt_main_table(n):
    T, digest, _ = main_table([OP_HALT], (), n)
    return T, digest


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
    by_col = {q: [(ev, {col for (r, is_main, col) in ev.inputs if not is_main and r == 1})
                  for ev in tran_evs if (1, False, q) in ev.inputs] for q in base_aux}
    unconstrained = set(unsolved0)
    for i in range(n - 1):
        solved = set()
        nxt = A[i + 1]
        progress = True
        while progress and len(solved) < 49:
            progress = False
            for q in base_aux:
                if q in solved: continue
                for ev, next_refs in by_col[q]:
                    others = next_refs - {q} - solved
                    u = _solve_affine(ev, rows_m[i], A[i], rows_m[i + 1], nxt, challenges, 1, q)
                    if u is not None and others:
                        # the combined per-instruction constraints mention other, still unknown, columns of the next
                        # row in the branches of the instructions that are NOT executing; accept the root only if it
                        # does not move when those unknowns do
                        for o in others: nxt[o] = (0x1234567 + o, 1, 2)
                        u2 = _solve_affine(ev, rows_m[i], A[i], rows_m[i + 1], nxt, challenges, 1, q)
                        for o in others: nxt[o] = (0, 0, 0)
                        if u2 != u: u = None
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

"""Trace generation: a restatement of the VM (all 46 instructions), of the master main table (fill + pad + derived
columns) and of the master auxiliary table (`extend`), so that AIR-SATISFYING instances exist in this repository — with
them the oracle verifier runs with `check_air=True` (the out-of-domain AIR / quotient identity, stark.rs:1469-1540), which
synthetic tables can never pass.

TEST INFRASTRUCTURE ONLY.  Follows, table by table:
  aet.rs:95-215 (program hashing trace, lookup multiplicities), vm.rs:246-268, 1113-1190 (initial state, processor row),
  table/master_table.rs:881-1004 (fill order, pad, derived columns), table/{program,processor,op_stack,ram,jump_stack,
  hash,cascade,lookup,u32}.rs (fill / pad / extend), triton-constraint-builder/src/substitutions.rs:128-330
  (derived-column fill), vm.rs:270-1100 (instruction semantics, helper variables), triton-isa/src/op_stack.rs:61-255
  (underflow I/O), table/u32.rs:100-290, table/ram.rs:64-262 (sections, Bezout coefficients).
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
        from . import stark
        _AIR = stark.air()                     # one build per process (≈ 8 s), shared with the prover / verifier
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


# ---- a VM (vm.rs:361-1100) -----------------------------------------------------------------------------------------
# All 46 instructions.
from .isa_words import OPCODES, HAS_ARG, assemble   # noqa: E402

_NAME = {v: k for k, v in OPCODES.items()}
SUPPORTED = set(OPCODES)
U32_MAX = (1 << 32) - 1


class Execution:
    """what AlgebraicExecutionTrace (aet.rs:41-91) records"""

    def __init__(self):
        self.rows, self.op_stack_entries, self.output = [], [], []
        self.sponge_rows = []            # (CI, round number, state)                         aet.rs:276-297
        self.hash_traces = []            # permutation traces of `hash`                      aet.rs:264-274
        self.lookup_traces = []          # every trace whose S-box look-ups are counted, in execution order
        self.u32_entries = {}            # (instruction name, lhs, rhs) -> multiplicity       aet.rs:346-348 (IndexMap)
        self.ram_calls = []              # (clk, is_write, pointer, value)                    ram.rs:38-60


def execute(words, public_input=(), secret_input=(), initial_ram=None, secret_digests=()):
    """VM::run with full tracing (vm.rs:361-1111) -> Execution"""
    ex = Execution()
    ex.digest = [int(v) for v in tip5.hash_varlen(words)]
    stack = list(reversed(ex.digest)) + [0] * 11                # OpStack::new: list index 0 = deepest element
    jump_stack, inp, sec, digests = [], list(public_input), list(secret_input), [list(d) for d in secret_digests]
    ram = dict(initial_ram or {})
    sponge = None
    ex.multiplicities = [0] * len(words)
    ip = clk = 0
    while True:
        if not 0 <= ip < len(words):
            raise ValueError("instruction pointer out of bounds")
        name = _NAME[words[ip]]
        if name not in SUPPORTED:
            raise ValueError(f"instruction {name} is not restated")
        arg = words[ip + 1] if name in HAS_ARG else None
        nxt_ip = ip + (2 if name in HAS_ARG else 1)
        nia = arg if arg is not None else (words[nxt_ip] if nxt_ip < len(words) else 1)      # vm.rs:1178-1189
        st = lambda i: stack[len(stack) - 1 - i]                                           # noqa: E731
        hv = [0] * 6                                                                        # vm.rs:270-345
        if name in ("pop", "divine", "pick", "place", "dup", "swap", "read_mem", "write_mem", "read_io", "write_io"):
            hv[:4] = [(arg >> k) & 1 for k in range(4)]
        elif name == "skiz":
            hv[0] = inv_or_zero(st(0))
            hv[1:6] = [nia % 2, (nia >> 1) % 4, (nia >> 3) % 4, (nia >> 5) % 4, nia >> 7]
        elif name == "recurse_or_return":
            hv[0] = inv_or_zero(st(6) - st(5))
        elif name == "split":
            if st(0) & U32_MAX:
                hv[0] = inv_or_zero((st(0) >> 32) - U32_MAX)
        elif name == "eq":
            hv[0] = inv_or_zero(st(1) - st(0))
        elif name == "sponge_absorb_mem":
            hv[:6] = [ram.get((st(0) + k) % P, 0) for k in range(4, 10)]
        elif name == "merkle_step":
            hv[:5] = digests[0] if digests else [0] * 5
            hv[5] = st(5) % 2
        elif name == "merkle_step_mem":
            hv[:5] = [ram.get((st(7) + k) % P, 0) for k in range(5)]
            hv[5] = st(5) % 2
        elif name == "b_horner_step":
            hv[0] = ram.get(st(5), 0)
        elif name == "x_horner_step":
            hv[2], hv[1], hv[0] = ram.get(st(5), 0), ram.get((st(5) - 1) % P, 0), ram.get((st(5) - 2) % P, 0)
        ex.rows.append(dict(clk=clk, ip=ip, ci=words[ip], nia=nia, jsp=len(jump_stack),
                            jso=jump_stack[-1][0] if jump_stack else 0, jsd=jump_stack[-1][1] if jump_stack else 0,
                            st=[st(i) for i in range(16)], osp=len(stack), hv=hv))
        ex.multiplicities[ip] += 1
        io = []                                                   # underflow IO of this instruction (op_stack.rs:61-103)

        def underflow():
            return stack[len(stack) - 17] if len(stack) > 16 else 0

        def push(e):
            stack.append(e % P)
            io.append(("w", underflow()))

        def pop():
            io.append(("r", underflow()))
            if len(stack) <= 16:
                raise ValueError("op stack too shallow")
            return stack.pop()

        def pop_u32():
            v = pop()
            if v > U32_MAX: raise ValueError("not a u32")
            return v

        def need_u32(i):
            if st(i) > U32_MAX: raise ValueError("not a u32")

        def u32_call(instr, lhs, rhs):
            key = (instr, lhs % P, rhs % P)
            ex.u32_entries[key] = ex.u32_entries.get(key, 0) + 1

        def permutation(state, sponge_ci=None):
            trace = tip5_trace(state)
            ex.lookup_traces.append(trace)
            if sponge_ci is None:
                ex.hash_traces.append(trace)
            else:
                ex.sponge_rows.extend((sponge_ci, rnd, row) for rnd, row in enumerate(trace))
            return trace

        def pop_x(): return (pop(), pop(), pop())

        def ram_read(ptr):
            v = ram.get(ptr % P, 0)
            ex.ram_calls.append((clk, 0, ptr % P, v))
            return v

        def merkle_step(sibling):                                   # vm.rs:1026-1062
            need_u32(5)
            node_index = st(5)
            acc = [pop() for _ in range(5)]
            left, right = (acc, sibling) if node_index % 2 == 0 else (sibling, acc)
            out = permutation(left + right + [1] * 6)[-1][:5]
            for e in reversed(out): push(e)
            stack[len(stack) - 1 - 5] = node_index // 2
            u32_call("split", node_index, node_index // 2)

        def horner_step(coefficient):                               # vm.rs:1064-1111
            x = (st(0), st(1), st(2))
            acc = F.xadd(F.xmul((st(7), st(8), st(9)), x), coefficient)
            for k in range(3): stack[len(stack) - 1 - (7 + k)] = acc[k]

        def push_x(x):
            for c in reversed(x): push(c)

        halting = False
        ip_after = nxt_ip
        if name == "halt": halting = True
        elif name == "nop": pass
        elif name == "push": push(arg)
        elif name == "pop":
            for _ in range(arg): pop()
        elif name == "divine":
            if len(sec) < arg: raise IndexError("secret input exhausted")
            for _ in range(arg): push(sec.pop(0))
        elif name == "pick":
            io.append(("r", underflow()))
            push(stack.pop(len(stack) - 1 - arg))
        elif name == "place":
            e = pop()
            stack.insert(len(stack) - arg, e)
            io.append(("w", underflow()))
        elif name == "dup": push(st(arg))
        elif name == "swap":
            i0, i1 = len(stack) - 1, len(stack) - 1 - arg
            stack[i0], stack[i1] = stack[i1], stack[i0]
        elif name == "skiz":
            if pop() == 0:
                ip_after = nxt_ip + (2 if _NAME[words[nxt_ip]] in HAS_ARG else 1)
        elif name == "call": jump_stack.append((ip + 2, arg)); ip_after = arg
        elif name == "return": ip_after = jump_stack.pop()[0]
        elif name == "recurse": ip_after = jump_stack[-1][1]
        elif name == "recurse_or_return":
            ip_after = jump_stack.pop()[0] if st(5) == st(6) else jump_stack[-1][1]
        elif name == "assert":
            if st(0) != 1: raise ValueError("assertion failed")
            pop()
        elif name == "read_mem":
            ptr = pop()
            for _ in range(arg):
                push(ram_read(ptr))
                ptr = (ptr - 1) % P
            push(ptr)
        elif name == "write_mem":
            ptr = pop()
            for _ in range(arg):
                v = pop()
                ex.ram_calls.append((clk, 1, ptr, v))
                ram[ptr] = v
                ptr = (ptr + 1) % P
            push(ptr)
        elif name == "hash":
            to_hash = [pop() for _ in range(10)]
            out = permutation(to_hash + [1] * 6)[-1][:5]            # Domain::FixedLength: capacity of ones
            for e in reversed(out): push(e)
        elif name == "assert_vector":
            if any(st(i) != st(i + 5) for i in range(5)): raise ValueError("vector assertion failed")
            for _ in range(5): pop()
        elif name == "sponge_init":
            sponge = [0] * 16
            ex.sponge_rows.append((OPCODES["sponge_init"], 0, list(sponge)))
        elif name == "sponge_absorb":
            if sponge is None: raise ValueError("sponge not initialized")
            sponge[:10] = [pop() for _ in range(10)]
            sponge = list(permutation(sponge, OPCODES["sponge_absorb"])[-1])
        elif name == "sponge_absorb_mem":                          # vm.rs:699-729
            if sponge is None: raise ValueError("sponge not initialized")
            ptr = pop()
            for i in range(10):
                sponge[i] = ram_read(ptr)
                ptr = (ptr + 1) % P
                if i < 4: stack[len(stack) - 1 - i] = sponge[i]
            push(ptr)
            sponge = list(permutation(sponge, OPCODES["sponge_absorb"])[-1])
        elif name == "merkle_step":
            if not digests: raise IndexError("secret digests exhausted")
            merkle_step(digests.pop(0))
        elif name == "merkle_step_mem":
            need_u32(5)
            ptr = st(7)
            sibling = []
            for _ in range(5):
                sibling.append(ram_read(ptr))
                ptr = (ptr + 1) % P
            stack[len(stack) - 1 - 7] = ptr
            merkle_step(sibling)
        elif name == "b_horner_step":
            ptr = st(5)
            c0 = ram_read(ptr)
            stack[len(stack) - 1 - 5] = (ptr - 1) % P
            horner_step((c0, 0, 0))
        elif name == "x_horner_step":
            ptr = st(5)
            c2 = ram_read(ptr); c1 = ram_read(ptr - 1); c0 = ram_read(ptr - 2)
            stack[len(stack) - 1 - 5] = (ptr - 3) % P
            horner_step((c0, c1, c2))
        elif name == "sponge_squeeze":
            if sponge is None: raise ValueError("sponge not initialized")
            for i in reversed(range(10)): push(sponge[i])
            sponge = list(permutation(sponge, OPCODES["sponge_squeeze"])[-1])
        elif name == "add": x = pop(); y = pop(); push(x + y)
        elif name == "addi": stack[-1] = (stack[-1] + arg) % P
        elif name == "mul": x = pop(); y = pop(); push(x * y)
        elif name == "invert":
            if st(0) == 0: raise ValueError("inverse of zero")
            push(F.inv(pop()))
        elif name == "eq": x = pop(); y = pop(); push(1 if x == y else 0)
        elif name == "split":
            top = pop()
            lo, hi = top & U32_MAX, top >> 32
            push(hi); push(lo)
            u32_call("split", lo, hi)
        elif name == "lt":
            need_u32(0); need_u32(1)
            lhs = pop_u32(); rhs = pop_u32(); push(1 if lhs < rhs else 0)
            u32_call("lt", lhs, rhs)
        elif name == "and":
            need_u32(0); need_u32(1)
            lhs = pop_u32(); rhs = pop_u32(); push(lhs & rhs)
            u32_call("and", lhs, rhs)
        elif name == "xor":
            need_u32(0); need_u32(1)
            lhs = pop_u32(); rhs = pop_u32(); push(lhs ^ rhs)
            u32_call("and", lhs, rhs)                              # a ^ b = a + b - 2 (a & b), vm.rs:860-865
        elif name == "log_2_floor":
            need_u32(0)
            if st(0) == 0: raise ValueError("logarithm of zero")
            top = pop_u32(); push(top.bit_length() - 1)
            u32_call("log_2_floor", top, 0)
        elif name == "pow":
            need_u32(1)
            base = pop(); exponent = pop_u32(); push(pow(base, exponent, P))
            u32_call("pow", base, exponent)
        elif name == "div_mod":
            need_u32(0); need_u32(1)
            if st(1) == 0: raise ValueError("division by zero")
            numerator = pop_u32(); denominator = pop_u32()
            quotient, remainder = divmod(numerator, denominator)
            push(quotient); push(remainder)
            u32_call("lt", remainder, denominator)
            u32_call("split", numerator, quotient)
        elif name == "pop_count":
            need_u32(0)
            top = pop_u32(); push(bin(top).count("1"))
            u32_call("pop_count", top, 0)
        elif name == "xx_add": x = pop_x(); y = pop_x(); push_x(F.xadd(x, y))
        elif name == "xx_mul": x = pop_x(); y = pop_x(); push_x(F.xmul(x, y))
        elif name == "x_invert":
            if (st(0), st(1), st(2)) == (0, 0, 0): raise ValueError("inverse of zero")
            push_x(F.xinv(pop_x()))
        elif name == "xb_mul": x = pop(); y = pop_x(); push_x(F.xmul((x, 0, 0), y))
        elif name == "read_io":
            if len(inp) < arg: raise IndexError("public input exhausted")
            for _ in range(arg): push(inp.pop(0))
        elif name == "write_io":
            for _ in range(arg): ex.output.append(pop())
        else:
            raise AssertionError(name)
        ip = ip_after
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
                ex.op_stack_entries.append((clk, 1 if t == "r" else 0, ptr, payload))
                if t == "w":
                    ptr += 1
        clk += 1
        if halting:
            break
    return ex


def run(words, public_input=(), secret_input=()):
    """-> (processor rows, op-stack table entries, instruction multiplicities, public output, program digest)"""
    ex = execute(words, public_input, secret_input)
    return ex.rows, ex.op_stack_entries, ex.multiplicities, ex.output, ex.digest


# ---- co-processor tables -------------------------------------------------------------------------------------------
def _program_hash(program):
    """aet.rs:150-190 -> (padded program, hash-table rows [(round, state)], permutation traces)"""
    padded_len = -(-(len(program) + 1) // 10) * 10
    padded_program = (list(program) + [1] + [0] * 10)[:padded_len]
    rows, traces = [], []
    sponge = [0] * 16
    for c0 in range(0, padded_len, 10):
        sponge[:10] = padded_program[c0:c0 + 10]
        trace = tip5_trace(sponge)
        traces.append(trace)
        rows += list(enumerate(trace))
        sponge = list(trace[-1])
    return padded_program, rows, traces, sponge[:5]


def _lookup_multiplicities(traces):
    """aet.rs:305-344: cascade multiplicities in first-use order (IndexMap), lookup-table multiplicities"""
    cascade_mult, lookup_mult = {}, [0] * 256
    for trace in traces:
        for row in trace[:-1]:
            for e in row[:4]:
                for limb in limbs16(e):
                    if limb in cascade_mult:
                        cascade_mult[limb] += 1
                    else:
                        cascade_mult[limb] = 1
                        lookup_mult[limb & 0xFF] += 1
                        lookup_mult[limb >> 8] += 1
    return cascade_mult, lookup_mult


def _u32_section(instr, lhs, rhs, multiplicity):
    """u32.rs:193-290: rows (CopyFlag, Bits, CI, LHS, RHS, Result, LookupMultiplicity) of one table section"""
    rows = []
    bits, flag, m = 0, 1, multiplicity
    while True:
        rows.append(dict(flag=flag, bits=bits, lhs=lhs, rhs=rhs, mult=m))
        if (lhs == 0 or instr == "pow") and rhs == 0:
            break
        lhs = lhs if instr == "pow" else lhs >> 1
        rhs >>= 1
        bits, flag, m = bits + 1, 0, 0
    last = rows[-1]
    last["result"] = {"split": 0, "lt": 2, "and": 0, "log_2_floor": P - 1, "pow": 1, "pop_count": 0}[instr]
    if instr == "lt" and last["bits"] == 0:
        last["result"] = 0
    for k in range(len(rows) - 2, -1, -1):
        row, nxt = rows[k], rows[k + 1]
        lsb_l, lsb_r, nr = row["lhs"] & 1, row["rhs"] & 1, nxt["result"]
        if instr == "split": r = nr
        elif instr == "lt":
            if nr in (0, 1): r = nr
            elif (lsb_l, lsb_r) == (0, 1): r = 1
            elif (lsb_l, lsb_r) == (1, 0): r = 0
            else: r = 0 if row["flag"] == 1 else 2
        elif instr == "and": r = (2 * nr + lsb_l * lsb_r) % P
        elif instr == "log_2_floor":
            r = P - 1 if row["lhs"] == 0 else (nr if nxt["lhs"] != 0 else row["bits"])
        elif instr == "pow": r = nr * nr % P if lsb_r == 0 else nr * nr % P * row["lhs"] % P
        else: r = (nr + lsb_l) % P
        row["result"] = r
    return rows


def _poly_mul(a, b):
    out = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                out[i + j] = (out[i + j] + x * y) % P
    return out


def bezout_coefficients(roots):
    """ram.rs:162-214: a, b with a*rp + b*fd = 1 for rp = prod (x - r), fd = rp' ; coefficient lists of length len(roots)"""
    n = len(roots)
    if n == 0:
        return [], []
    rp = [1]
    for r in roots:
        rp = _poly_mul(rp, [(-r) % P, 1])
    fd = [(k * rp[k]) % P for k in range(1, n + 1)]
    ev = lambda poly, x: sum(c * pow(x, k, P) for k, c in enumerate(poly)) % P        # noqa: E731
    b = [0] * n                                                    # Lagrange interpolation of 1 / fd(r) in the roots
    for r in roots:
        num, den = [1], 1
        for r2 in roots:
            if r2 != r:
                num = _poly_mul(num, [(-r2) % P, 1])
                den = den * (r - r2) % P
        scale = F.inv(ev(fd, r)) * F.inv(den) % P
        for k, c in enumerate(num):
            b[k] = (b[k] + c * scale) % P
    rem = _poly_mul(fd, b)                                         # a = (1 - fd b) / rp, an exact division
    rem = [(-c) % P for c in rem]
    rem[0] = (rem[0] + 1) % P
    a = [0] * max(1, len(rem) - n)
    for k in range(len(rem) - 1, n - 1, -1):                        # rp is monic of degree n
        q = rem[k]
        a[k - n] = q
        if q:
            for j in range(n + 1):
                rem[k - n + j] = (rem[k - n + j] - q * rp[j]) % P
    assert not any(rem), "division of 1 - fd*b by rp left a remainder"
    return (a + [0] * n)[:n], (b + [0] * n)[:n]


def table_heights(words, ex):
    """AlgebraicExecutionTrace::height_of_table (aet.rs:137-168)"""
    padded_len = -(-(len(words) + 1) // 10) * 10
    traces = _program_hash(words)[2] + ex.lookup_traces
    u32 = sum(1 if max(l if i != "pow" else 0, r) == 0 else 1 + max(l if i != "pow" else 0, r).bit_length()
              for (i, l, r) in ex.u32_entries)
    return dict(program=padded_len, processor=len(ex.rows), op_stack=len(ex.op_stack_entries), ram=len(ex.ram_calls),
                jump_stack=len(ex.rows), hash=6 * (padded_len // 10) + len(ex.sponge_rows) + 6 * len(ex.hash_traces),
                cascade=len(_lookup_multiplicities(traces)[0]), lookup=256, u32=u32)


def _hash_table_row(ci, rnd, st):
    """hash.rs:60-70 `trace_row_to_table_row` (+ CI, as aet.rs:203-205, 264-297 set it): one row of the AET's hash traces"""
    c = MAIN["hash"]
    names = c.names
    row = [0] * c.COUNT
    row[names.index("CI")], row[names.index("RoundNumber")] = ci, rnd
    for e in range(4):
        lb = limbs16(st[e])
        for k, part in enumerate(("Lowest", "MidLow", "MidHigh", "Highest")):
            row[names.index(f"State{e}{part}LkIn")] = lb[k]
            row[names.index(f"State{e}{part}LkOut")] = lookup16(lb[k])
        row[names.index(f"State{e}Inv")] = inv_or_zero((1 << 32) - 1 - ((lb[3] << 16) + lb[2]))
    for e in range(4, 16):
        row[names.index(f"State{e}")] = st[e]
    for k in range(16):
        row[names.index(f"Constant{k}")] = tip5.ROUND_CONSTANTS[16 * rnd + k] if rnd < 5 else 0
    return row


def aet_arrays(words, ex):
    """The arrays an AlgebraicExecutionTrace holds (aet.rs:41-91), in the reference's layout: what crosses the C ABI as
    `tvm_aet` (include/tvm_b200.h).  `ex` = execute(words, ...)."""
    program = list(words)
    h_op = OPCODES["hash"]
    _, program_hash_rows, program_traces, _ = _program_hash(program)
    cascade_mult, lookup_mult = _lookup_multiplicities(program_traces + ex.lookup_traces)
    proc = np.zeros((len(ex.rows), 39), dtype=np.uint64)
    for i, r in enumerate(ex.rows):
        ci = r["ci"]
        proc[i] = ([r["clk"], 0, r["ip"], ci, r["nia"]] + [(ci >> b) & 1 for b in range(7)] + [r["jsp"], r["jso"], r["jsd"]]
                   + list(r["st"]) + [r["osp"]] + list(r["hv"]) + [0])

    def arr(rows, width):
        return np.array(rows, dtype=np.uint64).reshape(-1, width)
    return dict(
        program=np.array(program, dtype=np.uint64),
        instruction_multiplicities=np.array(ex.multiplicities, dtype=np.uint32),
        processor_trace=proc,
        op_stack_underflow_trace=arr([[clk, shrink, ptr, payload] for clk, shrink, ptr, payload in ex.op_stack_entries], 4),
        ram_trace=arr([[clk, 0 if is_write else 1, ptr, val, 0, 0, 0] for clk, is_write, ptr, val in ex.ram_calls], 7),
        program_hash_trace=arr([_hash_table_row(h_op, rnd, st) for rnd, st in program_hash_rows], 67),
        sponge_trace=arr([_hash_table_row(ci, rnd, st) for ci, rnd, st in ex.sponge_rows], 67),
        hash_trace=arr([_hash_table_row(h_op, rnd, st) for trace in ex.hash_traces for rnd, st in enumerate(trace)], 67),
        u32_entries=arr([[OPCODES[i], lhs, rhs, mult] for (i, lhs, rhs), mult in ex.u32_entries.items()], 4),
        cascade_table_lookup_multiplicities=arr([[limb, m] for limb, m in cascade_mult.items()], 2),
        lookup_table_lookup_multiplicities=np.array(lookup_mult, dtype=np.uint64))


# ---- main table -------------------------------------------------------------------------------------
def main_table(words, public_input, n, secret_input=(), initial_ram=None, secret_digests=(), evaluate_substitutions=False):
    """[379][n] canonical ints: MasterMainTable::new + pad (master_table.rs:881-1004).
    -> (table, program digest, public output).  The 230 degree-lowering columns come from the generated straight-line
    rules (oracle/c/aux_extend.c) unless `evaluate_substitutions`: then the substitution circuits are evaluated node by
    node in Python (`fill_derived_main_columns`, ~20x slower; tests/test_aux_extend.py checks the two against each other)."""
    assert n >= 256 and n & (n - 1) == 0
    T = np.zeros((NUM_MAIN, n), dtype=np.uint64)         # every entry is a canonical field element < p < 2^64
    program = list(words)
    ex = execute(program, public_input, secret_input, initial_ram, secret_digests)
    rows, plen = ex.rows, len(ex.rows)
    heights = table_heights(program, ex)
    assert max(heights.values()) <= n, f"table heights {heights} exceed {n}"

    padded_program, program_hash_rows, program_traces, digest = _program_hash(program)
    assert digest == ex.digest
    padded_len = len(padded_program)
    cascade_mult, lookup_mult = _lookup_multiplicities(program_traces + ex.lookup_traces)
    h_op = OPCODES["hash"]
    hash_rows = ([(1, h_op, rnd, st) for rnd, st in program_hash_rows]                       # hash.rs:241-268
                 + [(2, ci, rnd, st) for ci, rnd, st in ex.sponge_rows]
                 + [(3, h_op, rnd, st) for trace in ex.hash_traces for rnd, st in enumerate(trace)])

    # -- program table (program.rs:33-113)
    c = MAIN["program"]
    idx = np.arange(n, dtype=np.uint64)
    T[c.Address] = idx
    T[c.IndexInChunk] = idx % 10
    T[c.MaxMinusIndexInChunkInv] = np.array([inv_or_zero(9 - k) for k in range(10)], dtype=np.uint64)[idx % 10]
    T[c.Instruction, :padded_len] = padded_program
    T[c.LookupMultiplicity, :len(program)] = ex.multiplicities
    T[c.IsHashInputPadding, len(program):] = 1
    T[c.IsTablePadding, padded_len:] = 1

    # -- op stack table (op_stack.rs:179-211): sorted by (stack pointer, clk); padding copies the last row
    c = MAIN["op_stack"]
    os_sorted = sorted(ex.op_stack_entries, key=lambda e: (e[2], e[0]))
    clk_jump_diffs = []
    for i, (clk, shrink, ptr, payload) in enumerate(os_sorted):
        T[c.CLK, i], T[c.IB1ShrinkStack, i], T[c.StackPointer, i], T[c.FirstUnderflowElement, i] = clk, shrink, ptr, payload
        if i and os_sorted[i - 1][2] == ptr:
            clk_jump_diffs.append(clk - os_sorted[i - 1][0])
    if os_sorted:
        last = len(os_sorted) - 1
        for col in (c.CLK, c.StackPointer, c.FirstUnderflowElement):
            T[col, last + 1:] = T[col, last]
        T[c.IB1ShrinkStack, last + 1:] = 2
    else:
        T[c.IB1ShrinkStack, :] = 2
        T[c.StackPointer, :] = 16

    # -- RAM table (ram.rs:64-141, 216-262): sorted by (pointer, clk); Bezout coefficients change with the pointer
    c = MAIN["ram"]
    ram_sorted = sorted(ex.ram_calls, key=lambda e: (e[2], e[0]))
    if ram_sorted:
        unique = list(dict.fromkeys(e[2] for e in ram_sorted))
        bez0, bez1 = bezout_coefficients(unique)
        cur0, cur1 = bez0.pop(), bez1.pop()
        for i, (clk, is_write, ptr, val) in enumerate(ram_sorted):
            if i:
                prev = ram_sorted[i - 1]
                if prev[2] == ptr:
                    clk_jump_diffs.append(clk - prev[0])
                else:
                    cur0, cur1 = bez0.pop(), bez1.pop()
                T[c.InverseOfRampDifference, i - 1] = inv_or_zero(ptr - prev[2])
            T[c.CLK, i], T[c.InstructionType, i], T[c.RamPointer, i], T[c.RamValue, i] = clk, 0 if is_write else 1, ptr, val
            T[c.BezoutCoefficientPolynomialCoefficient0, i], T[c.BezoutCoefficientPolynomialCoefficient1, i] = cur0, cur1
        assert not bez0 and not bez1
        last = len(ram_sorted) - 1
        for col in range(c.start, c.start + c.COUNT):
            T[col, last + 1:] = T[col, last]
        T[c.InstructionType, last + 1:] = 2
    else:
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
    head, tail = np.array(js[:k_max + 1], dtype=np.uint64).reshape(-1, 5), np.array(js[k_max + 1:], dtype=np.uint64).reshape(-1, 5)
    npad = n - plen
    for k, col in enumerate((c.CLK, c.CI, c.JSP, c.JSO, c.JSD)):
        T[col, :k_max + 1] = head[:, k]
        T[col, k_max + 1:k_max + 1 + npad] = np.arange(plen, n, dtype=np.uint64) if k == 0 else head[k_max, k]
        T[col, k_max + 1 + npad:] = tail[:, k]

    # -- processor table (vm.rs:1113-1190, processor.rs:45-95)
    c = MAIN["processor"]

    def column(values, col):                               # executed rows, then the last row repeated as padding
        T[col, :plen] = values
        T[col, plen:] = values[-1]
    for key, col in (("clk", c.CLK), ("ip", c.IP), ("ci", c.CI), ("nia", c.NIA), ("jsp", c.JSP), ("jso", c.JSO), ("jsd", c.JSD),
                     ("osp", c.OpStackPointer)):
        column(np.array([r[key] for r in rows], dtype=np.uint64), col)
    T[c.CLK, plen:] = np.arange(plen, n, dtype=np.uint64)
    ci = T[c.CI]
    for b in range(7):
        T[c.IB0 + b] = (ci >> np.uint64(b)) & np.uint64(1)
    st_cols = np.array([r["st"] for r in rows], dtype=np.uint64)
    hv_cols = np.array([r["hv"] for r in rows], dtype=np.uint64)
    for k in range(16):
        column(st_cols[:, k], c.ST0 + k)
    for k in range(6):
        column(hv_cols[:, k], c.HV0 + k)
    T[c.IsPadding, plen:] = 1
    np.add.at(T[c.ClockJumpDifferenceLookupMultiplicity], np.array(clk_jump_diffs, dtype=np.int64), np.uint64(1))
    if n > plen:
        T[c.ClockJumpDifferenceLookupMultiplicity, 1] = (int(T[c.ClockJumpDifferenceLookupMultiplicity, 1]) + (n - plen)) % P

    # -- hash table (hash.rs:36-302)
    c = MAIN["hash"]
    names = c.names

    def col(name): return c.start + names.index(name)
    parts = ("Lowest", "MidLow", "MidHigh", "Highest")
    for i, (mode, ci, rnd, st) in enumerate(hash_rows):
        T[c.Mode, i] = mode
        T[c.CI, i] = ci
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
    nh = len(hash_rows)
    for e in range(4):
        T[col(f"State{e}Inv"), nh:] = zero_inv
    for k in range(16):
        T[col(f"Constant{k}"), nh:] = tip5.ROUND_CONSTANTS[k]
    T[c.Mode, nh:] = 0
    T[c.CI, nh:] = h_op

    # -- cascade table (cascade.rs:41-66): first-use order of the multiplicity map
    c = MAIN["cascade"]
    for i, (limb, m) in enumerate(cascade_mult.items()):
        T[c.LookInLo, i] = limb & 0xFF
        T[c.LookInHi, i] = limb >> 8
        T[c.LookOutLo, i] = lookup8(limb & 0xFF)
        T[c.LookOutHi, i] = lookup8(limb >> 8)
        T[c.LookupMultiplicity, i] = m
    T[c.IsPadding, len(cascade_mult):] = 1

    # -- lookup table (lookup.rs:84-116)
    c = MAIN["lookup"]
    for i in range(256):
        T[c.LookIn, i] = i
        T[c.LookOut, i] = lookup8(i)
        T[c.LookupMultiplicity, i] = lookup_mult[i]
    T[c.IsPadding, 256:] = 1

    # -- u32 table (u32.rs:100-154, 193-290): one section per distinct (instruction, operands), first-use order
    c = MAIN["u32"]
    i = 0
    for (instr, lhs, rhs), mult in ex.u32_entries.items():
        for r in _u32_section(instr, lhs, rhs, mult):
            T[c.CopyFlag, i], T[c.Bits, i], T[c.CI, i] = r["flag"], r["bits"], OPCODES[instr]
            T[c.BitsMinus33Inv, i] = F.inv((r["bits"] - 33) % P)
            T[c.LHS, i], T[c.RHS, i] = r["lhs"], r["rhs"]
            T[c.LhsInv, i], T[c.RhsInv, i] = inv_or_zero(r["lhs"]), inv_or_zero(r["rhs"])
            T[c.Result, i], T[c.LookupMultiplicity, i] = r["result"], r["mult"]
            i += 1
    pad_ci, pad_lhs, pad_lhs_inv, pad_result = OP_SPLIT, 0, 0, 0
    if i:
        pad_ci, pad_lhs, pad_lhs_inv, pad_result = T[c.CI, i - 1], T[c.LHS, i - 1], T[c.LhsInv, i - 1], T[c.Result, i - 1]
        if pad_ci == OPCODES["lt"]:
            pad_result = 2
    T[c.CI, i:], T[c.LHS, i:], T[c.LhsInv, i:], T[c.Result, i:] = pad_ci, pad_lhs, pad_lhs_inv, pad_result
    T[c.BitsMinus33Inv, i:] = F.inv((-33) % P)

    if evaluate_substitutions:
        fill_derived_main_columns(T)
    else:
        from . import corc
        T[149:] = corc.fill_derived_main(T)[149:]
    return T, ex.digest, ex.output


def padded_height(words, public_input=(), secret_input=(), initial_ram=None, secret_digests=()):
    """AlgebraicExecutionTrace::padded_height (aet.rs:99-135)"""
    h = max(table_heights(list(words), execute(list(words), public_input, secret_input, initial_ram, secret_digests)).values())
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

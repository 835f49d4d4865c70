"""Generate the auxiliary-table extension (MasterMainTable::extend, master_table.rs:1006-1075) FROM THE AIR.

Every one of the 49 auxiliary base columns is a running product / running evaluation / running sum of logarithmic
derivatives (cross_table_argument.rs).  The reference fills them with one hand-written sequential loop per column
(processor.rs:139-640, op_stack.rs:213-290, ram.rs:105-255, jump_stack.rs:207-280, hash.rs:304-460, cascade.rs:68-130,
lookup.rs:118-180, u32.rs:156-230, program.rs:115-190).  The AIR pins the very same recurrences: for column q the
initial constraint mentioning q is affine in aux[0][q], and each transition constraint mentioning next[q] is affine in
(cur[q], next[q]):

        C = alpha * next[q] + beta * cur[q] + gamma          alpha, beta, gamma: polynomials in the two main rows,
                                                             the challenges and auxiliary columns of LOWER level

so   next[q] = a * cur[q] + b   with  a = -beta/alpha, b = -gamma/alpha  — an affine map per row.  A column is the
inclusive scan of its row maps under composition, which is what the device does (csrc/aux_extend.cu); alpha, beta,
gamma are obtained here by symbolic differentiation of the constraint circuits, so the generated code cannot drift
from the constraints the quotient kernels enforce.

Rules per column: several constraints can mention next[q] (case distinctions); the first whose alpha is non-zero on a
row defines the row's map, if none does the column carries over.  The processor's instruction-specific constraints
(processor.rs:363-434: sum over instructions of deselector * constraint) are taken per instruction, selected by the
current instruction's opcode and the next row's padding flag, instead of through the 7-bit deselector polynomials.

The same text is emitted for the CPU oracle's checker (oracle/c/aux_extend_gen.inc) and for the device
(csrc/aux_gen/aux_extend_gen.inc); the including file supplies the accessor macros.

Run:  python -m airgen.extend_gen
"""
import os

from .build import PROVIDERS, CATEGORIES, _FN, build_air
from .circuit import Builder, P, reachable_postorder
from .columns import Env, NUM_AUX_COLUMNS
from .isa import ALL_INSTRUCTIONS, OPCODE
from .tables import processor

R = (1 << 64) % P
NUM_BASE_AUX = NUM_AUX_COLUMNS                # 49 table columns; then 41 derived columns and the batch randomizer (column 90)
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUTS = [os.path.join(ROOT, "oracle", "c", "aux_extend_gen.inc"),
        os.path.join(ROOT, "triton-vm_b200", "csrc", "aux_gen", "aux_extend_gen.inc")]
NUM_PROCESSOR_LINKING = 7                      # processor.rs:338-346


def mont(v): return v * R % P


# ---- circuit transforms ----------------------------------------------------------------------------------------------
class Xform:
    """copies circuits into builder `nb` (constant folding and structural sharing come with it)"""

    def __init__(self, nb):
        self.nb = nb
        self._copy = {}

    def leaf(self, n, mapping=None):
        nb = self.nb
        if n.kind == "B": return nb.b_constant(n.val)
        if n.kind == "X": return nb.x_constant(n.val)
        if n.kind == "C": return nb.challenge(n.val)
        if mapping and n.val in mapping: return mapping[n.val]
        return nb.input(*n.val)

    def copy(self, root, mapping=None):
        memo = self._copy if not mapping else {}
        for n in reachable_postorder([root]):
            if id(n) in memo: continue
            if n.kind in "+*":
                l, r = memo[id(n.lhs)], memo[id(n.rhs)]
                memo[id(n)] = (l + r) if n.kind == "+" else (l * r)
            else:
                memo[id(n)] = self.leaf(n, mapping)
        return memo[id(root)]

    def deriv(self, root, var):
        """d root / d var  (var = an input key)"""
        zero, one = self.nb.b_constant(0), self.nb.b_constant(1)
        d = {}
        for n in reachable_postorder([root]):
            if n.kind == "I":
                d[id(n)] = one if n.val == var else zero
            elif n.kind == "+":
                d[id(n)] = d[id(n.lhs)] + d[id(n.rhs)]
            elif n.kind == "*":
                dl, dr = d[id(n.lhs)], d[id(n.rhs)]
                acc = zero
                if not dl.n.is_zero(): acc = acc + dl * self.copy(n.rhs)
                if not dr.n.is_zero(): acc = acc + self.copy(n.lhs) * dr
                d[id(n)] = acc
            else:
                d[id(n)] = zero
        return d[id(root)]


def inputs_of(node):
    return {n.val for n in reachable_postorder([node]) if n.kind == "I"}


class Rule:
    """alpha * next + beta * cur + gamma == 0 (transition) / alpha * x + gamma == 0 (initial); guard: None or
    (opcode or None, next_is_padding)"""

    def __init__(self, alpha, beta, gamma, guard=None):
        self.alpha, self.beta, self.gamma, self.guard = alpha, beta, gamma, guard

    def nodes(self):
        return [m.n for m in (self.alpha, self.beta, self.gamma) if m is not None]

    def aux_deps(self, q):
        out = set()
        for nd in self.nodes():
            out |= {col for (row, is_main, col) in inputs_of(nd) if not is_main and col != q}
        return out


def make_rules(xf, constraint, q, dual, guard=None):
    zero = xf.nb.b_constant(0)
    if dual:
        nxt, cur = (1, False, q), (0, False, q)
        if nxt not in inputs_of(constraint): return None
        alpha, beta = xf.deriv(constraint, nxt), xf.deriv(constraint, cur)
        gamma = xf.copy(constraint, {nxt: zero, cur: zero})
    else:
        cur = (0, False, q)
        if cur not in inputs_of(constraint): return None
        alpha, beta = xf.deriv(constraint, cur), None
        gamma = xf.copy(constraint, {cur: zero})
    if alpha.n.is_zero(): return None
    for m in (alpha, beta, gamma):
        if m is None: continue
        own = {(row, is_main, col) for (row, is_main, col) in inputs_of(m.n) if not is_main and col == q}
        assert not own, f"constraint is not affine in auxiliary column {q}"
    return Rule(alpha, beta, gamma, guard)


def collect():
    """-> (init_rules[q], tran_rules[q]).  A constraint that mentions one auxiliary column of the next row (row 0 for
    the initial constraints) defines that column; one that mentions several (e.g. program.rs:283-296, where the send-chunk
    evaluation absorbs the next row's prepare-chunk evaluation) defines the only one of them no other constraint defines."""
    ixf, txf = Xform(Builder(dual=False)), Xform(Builder(dual=True))
    src_i, src_t = Builder(dual=False), Builder(dual=True)
    init_cs, tran_cs = [], []                                # (constraint node, guard)
    for prov in PROVIDERS:
        init_cs += [(m.n, None) for m in getattr(prov, _FN["init"])(src_i)]
        tran = [m.n for m in getattr(prov, _FN["tran"])(src_t)]
        if prov is processor:
            e = Env(src_t)
            for instr in ALL_INSTRUCTIONS:
                tran_cs += [(m.n, (OPCODE[instr], 0)) for m in processor.transition_constraints_for_instruction(e, instr)]
            tran_cs += [(m.n, (None, 1)) for m in processor.padding_row_constraints(e)]
            tran = tran[:3] + tran[-NUM_PROCESSOR_LINKING:]
        tran_cs += [(c, None) for c in tran]

    def assign(cs, xf, dual):
        row = 1 if dual else 0
        sets = [sorted({col for (r, is_main, col) in inputs_of(c) if not is_main and r == row and col < NUM_BASE_AUX})
                for c, _ in cs]
        defined = {s[0] for s in sets if len(s) == 1}
        owner = [s[0] if len(s) == 1 else None for s in sets]
        changed = True
        while changed:
            changed = False
            for k, s in enumerate(sets):
                if len(s) < 2 or owner[k] is not None: continue
                cand = [q for q in s if q not in defined]
                if len(cand) == 1:
                    owner[k] = cand[0]; defined.add(cand[0]); changed = True
                elif not cand:                    # all defined elsewhere: a consistency link, e.g. of another instruction
                    owner[k] = -1
        assert all(o is not None for o, s in zip(owner, sets) if s), "could not attribute a multi-column constraint"
        rules = {q: [] for q in range(NUM_BASE_AUX)}
        for (c, guard), q in zip(cs, owner):
            if q is None or q < 0: continue
            r = make_rules(xf, c, q, dual, guard)
            if r: rules[q].append(r)
        return rules

    return assign(init_cs, ixf, False), assign(tran_cs, txf, True)


def levels(init_rules, tran_rules):
    deps = {q: set() for q in range(NUM_BASE_AUX)}
    for q in range(NUM_BASE_AUX):
        for r in init_rules[q] + tran_rules[q]:
            deps[q] |= r.aux_deps(q)
    level = {}
    while len(level) < NUM_BASE_AUX:
        ready = [q for q in range(NUM_BASE_AUX) if q not in level and all(d in level for d in deps[q])]
        assert ready, f"cyclic dependency among auxiliary columns: { {q: sorted(deps[q]) for q in deps if q not in level} }"
        lv = 1 + max((level[d] for q in ready for d in deps[q]), default=-1)
        # columns become ready in waves; one wave = one level
        for q in ready:
            level[q] = max((level[d] + 1 for d in deps[q]), default=0)
        del lv
    return level, deps


# ---- emission --------------------------------------------------------------------------------------------------------
class Emitter:
    """straight-line typed code for a set of circuit nodes; shares temporaries inside one block"""

    def __init__(self, indent="    ", main_is_x=False):
        self.lines, self.name, self.isx, self.cnt, self.indent = [], {}, {}, 0, indent
        self.main_is_x = main_is_x            # the verifier evaluates the AIR on out-of-domain (X-field) main rows

    def tmp(self):
        self.cnt += 1
        return f"t{self.cnt}"

    def emit(self, roots):
        name, isx, lines = self.name, self.isx, self.lines
        for n in reachable_postorder(roots):
            key = id(n)
            if key in name: continue
            k = n.kind
            if k == "B":
                name[key] = f"0x{mont(n.val):016x}ULL"; isx[key] = False
            elif k == "X":
                v = self.tmp()
                lines.append(f"const xfe {v} = {{0x{mont(n.val[0]):016x}ULL, 0x{mont(n.val[1]):016x}ULL, 0x{mont(n.val[2]):016x}ULL}};")
                name[key] = v; isx[key] = True
            elif k == "C":
                v = self.tmp(); lines.append(f"const xfe {v} = CH({n.val});"); name[key] = v; isx[key] = True
            elif k == "I":
                row, is_main, col = n.val
                v = self.tmp()
                if is_main:
                    ty = "xfe" if self.main_is_x else "u64"
                    lines.append(f"const {ty} {v} = {'MN' if row else 'MC'}({col});"); isx[key] = self.main_is_x
                else:
                    lines.append(f"const xfe {v} = {'AN' if row else 'AC'}({col});"); isx[key] = True
                name[key] = v
            else:
                lx, rx = isx[id(n.lhs)], isx[id(n.rhs)]
                ln, rn = name[id(n.lhs)], name[id(n.rhs)]
                v = self.tmp()
                if k == "*":
                    if lx and rx: e, x = f"xmul({ln}, {rn})", True
                    elif lx: e, x = f"xmulb({ln}, {rn})", True
                    elif rx: e, x = f"xmulb({rn}, {ln})", True
                    else: e, x = f"fmul({ln}, {rn})", False
                else:
                    if lx and rx: e, x = f"xadd({ln}, {rn})", True
                    elif lx: e, x = f"xaddb({ln}, {rn})", True
                    elif rx: e, x = f"xaddb({rn}, {ln})", True
                    else: e, x = f"fadd({ln}, {rn})", False
                lines.append(f"const {'xfe' if x else 'u64'} {v} = {e};")
                name[key] = v; isx[key] = x

    def ref_x(self, m):
        """expression of X-field type for node m"""
        nm = self.name[id(m.n)]
        return nm if self.isx[id(m.n)] else f"xlift({nm})"

    def take(self):
        out = [self.indent + l for l in self.lines]
        self.lines = []
        return out


def emit_rule(r, dual, ind="    "):
    """block that returns from the enclosing function when the rule applies"""
    em = Emitter(ind + "  ")
    out = []
    guard = []
    if r.guard:
        op, pad = r.guard
        if op is not None: guard.append(f"MC({processor.C.CI}) == 0x{mont(op):016x}ULL")
        guard.append(f"MN({processor.C.IsPadding}) == 0x{mont(pad):016x}ULL")
    out.append(f"{ind}{'if (' + ' && '.join(guard) + ') ' if guard else ''}{{")
    unit = r.alpha.n.is_one()
    if not unit:
        em.emit([r.alpha.n])
        out += em.take()
        out.append(f"{ind}  const xfe alpha = {em.ref_x(r.alpha)};")
        out.append(f"{ind}  if (!xis_zero(alpha)) {{")
        em.indent = ind + "    "
    rest = [m.n for m in (r.beta, r.gamma) if m is not None]
    em.emit(rest)
    out += em.take()
    i2 = ind + ("    " if not unit else "  ")
    if not unit:
        out.append(f"{i2}const xfe ninv = xneg(xinv(alpha));")
        if dual: out.append(f"{i2}*a = xmul({em.ref_x(r.beta)}, ninv);")
        out.append(f"{i2}*b = xmul({em.ref_x(r.gamma)}, ninv);")
    else:
        if dual: out.append(f"{i2}*a = xneg({em.ref_x(r.beta)});")
        out.append(f"{i2}*b = xneg({em.ref_x(r.gamma)});")
    out.append(f"{i2}return 1;")
    if not unit:
        out.append(f"{ind}  }}")
    out.append(f"{ind}}}")
    return out


def generate():
    init_rules, tran_rules = collect()
    level, deps = levels(init_rules, tran_rules)
    nlev = 1 + max(level.values())
    L = ["/* GENERATED by triton-vm_b200/airgen/extend_gen.py — do not edit.",
         " * Auxiliary-table extension derived from the AIR: per auxiliary base column q the row map",
         " *     aux[i][q] = a * aux[i-1][q] + b",
         " * (initial value for i = 0).  Including file defines AUXGEN_FN, AUXGEN_ARGS, MC/MN (main row i-1 / i, Montgomery),",
         " * AC/AN (auxiliary row i-1 / i as xfe), CH (challenge as xfe), AW (store into the current row), AUXGEN_PASS (the",
         " * argument forwarded to callees), AUXGEN_TOUCH (statement marking the argument used), MW (store a base-field value",
         " * into the current row of the main table) and the field ops. */",
         f"#define AUXGEN_NUM_BASE {NUM_BASE_AUX}",
         f"#define AUXGEN_NUM_LEVELS {nlev}",
         "static const int AUXGEN_LEVEL[AUXGEN_NUM_BASE] = {" + ", ".join(str(level[q]) for q in range(NUM_BASE_AUX)) + "};", ""]
    stats = []
    for q in range(NUM_BASE_AUX):
        # initial value: single-row builder, inputs are row 0 -> MC / AC
        L.append(f"AUXGEN_FN int auxgen_init_{q}(AUXGEN_ARGS, xfe *b) {{")
        L.append("    AUXGEN_TOUCH;")
        for r in init_rules[q]:
            L += emit_rule(r, False)
        L += ["    (void)b; return 0;", "}", ""]
        L.append(f"AUXGEN_FN int auxgen_tran_{q}(AUXGEN_ARGS, xfe *a, xfe *b) {{")
        L.append("    AUXGEN_TOUCH;")
        for r in tran_rules[q]:
            L += emit_rule(r, True)
        L += ["    (void)a; (void)b; return 0;", "}", ""]
        stats.append((q, level[q], len(init_rules[q]), len(tran_rules[q]), sorted(deps[q])))
    for kind, extra in (("init", "xfe *b"), ("tran", "xfe *a, xfe *b")):
        L.append(f"AUXGEN_FN int auxgen_{kind}(int q, AUXGEN_ARGS, {extra}) {{")
        L.append("    switch (q) {")
        for q in range(NUM_BASE_AUX):
            L.append(f"    case {q}: return auxgen_{kind}_{q}(AUXGEN_PASS, {'b' if kind == 'init' else 'a, b'});")
        L += ["    default: return 0;", "    }", "}", ""]

    # derived (degree-lowering) auxiliary columns, substitutions.rs:163-330: value = -(rule with its own column := 0)
    air = build_air()
    for cat in CATEGORIES:
        rules = air.aux_subst[cat]
        start = air.subst_col_start[cat][1]
        dual = cat == "tran"
        nb = Builder(dual=dual)
        xf = Xform(nb)
        L.append(f"#define AUXGEN_NUM_DERIVED_{cat.upper()} {len(rules)}")
        L.append(f"#define AUXGEN_DERIVED_START_{cat.upper()} {start}")
        if not rules: continue                       # today only the transition constraints need lowering columns
        L.append(f"AUXGEN_FN void auxgen_derived_{cat}(AUXGEN_ARGS) {{")
        for k, rule in enumerate(rules):
            own = (0, False, start + k)
            assert own in inputs_of(rule)
            expr = xf.copy(rule, {own: nb.b_constant(0)})
            later = {col for (row, is_main, col) in inputs_of(expr.n) if not is_main and col >= start + k and row == 0}
            assert not later, "derived column depends on a later derived column of its own row"
            nxt_derived = {col for (row, is_main, col) in inputs_of(expr.n) if not is_main and col >= NUM_BASE_AUX and row == 1}
            assert not nxt_derived, "derived column reads a derived column of the next row"      # substitutions.rs:360-361
            em = Emitter("      ")
            em.emit([expr.n])
            L.append("    {")
            L += em.take()
            L.append(f"      AW({start + k}, xneg({em.ref_x(expr)}));")
            L.append("    }")
        L += ["}", ""]
    # derived (degree-lowering) MAIN columns, substitutions.rs:128-161, 237-300: sections init | cons (single row, every
    # row) and tran (rows i, i+1; the last row stays 0).  Base-field valued; a rule may read earlier derived columns of its
    # own row and, in the transition section, the next row's columns of the earlier sections.
    L.append(f"#define AUXGEN_NUM_BASE_MAIN {air.subst_col_start['init'][0]}")
    for cat in CATEGORIES:
        rules = air.main_subst[cat]
        start = air.subst_col_start[cat][0]
        dual = cat == "tran"
        nb = Builder(dual=dual)
        xf = Xform(nb)
        L.append(f"#define AUXGEN_NUM_DERIVED_MAIN_{cat.upper()} {len(rules)}")
        L.append(f"#define AUXGEN_DERIVED_MAIN_START_{cat.upper()} {start}")
        if not rules: continue
        L.append(f"AUXGEN_FN void auxgen_derived_main_{cat}(AUXGEN_ARGS) {{")
        L.append("    AUXGEN_TOUCH;")
        for k, rule in enumerate(rules):
            own = (0, True, start + k)
            assert own in inputs_of(rule)
            expr = xf.copy(rule, {own: nb.b_constant(0)})
            ins = inputs_of(expr.n)
            assert all(is_main for (_, is_main, _) in ins), "main substitution reads an auxiliary column"
            assert not {c for (row, _, c) in ins if row == 0 and c >= start + k}, "reads a later derived column of its row"
            assert not {c for (row, _, c) in ins if row == 1 and c >= start}, "reads the next row's column of its own section"
            em = Emitter("      ")
            em.emit([expr.n])
            assert not em.isx[id(expr.n)]
            L.append("    {")
            L += em.take()
            L.append(f"      MW({start + k}, fneg({em.name[id(expr.n)]}));")
            L.append("    }")
        L += ["}", ""]
    text = "\n".join(L) + "\n"
    for o in OUTS:
        os.makedirs(os.path.dirname(o), exist_ok=True)
        with open(o, "w") as f:
            f.write(text)
    return stats, nlev, len(text)


if __name__ == "__main__":
    stats, nlev, size = generate()
    for q, lv, ni, nt, dp in stats:
        print(f"aux {q:2d}: level {lv}, {ni} initial rule(s), {nt} transition rule(s), depends on {dp}")
    print("levels:", nlev, "generated bytes:", size)

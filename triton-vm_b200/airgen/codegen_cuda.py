"""Emit sm_100a CUDA source for the AIR quotient kernels from the lowered constraint circuits.

Reference behaviour being replaced: `all_quotients_combined` (triton-vm/src/table/
master_table.rs:1264-1363) calling the build-time generated `Evaluable::evaluate_*_constraints`
(triton-constraint-builder/src/codegen.rs:59-269).  For every quotient-domain row the result is
    sum_cat  zerofier_inverse_cat(row) * sum_j  weight_j * constraint_j(row, next_row)
with constraint order init | cons | tran | term and, inside a category, base-field-valued
constraints before extension-field-valued ones (codegen.rs:210-212).

B200 mapping: one thread per row, straight-line code with common sub-expressions shared inside a
chunk.  The 604 constraints are split into many small kernels ("chunks"): ncu showed the first
design (14 chunks of ~30k SASS instructions each, profiles/r01a_ncu_air_chunk06.md) issuing at
15 % with `stalled_no_instruction` dominating - straight-line code that large streams through the
32 KB L1.5 instruction cache once per warp.  A chunk is therefore sized to stay instruction-cache
resident (budget below, ~25 SASS instructions per cost unit); a node needed by two chunks is
recomputed and a chunk re-reads the table columns it touches, which HBM has ample headroom for.
Every chunk adds its zerofier-weighted partial sum into the row-planar X-field output.  The
weighted sum  sum_j w_j * c_j  is accumulated UNREDUCED (three 128-bit accumulators + overflow
words, one Montgomery reduction per chunk instead of one per product); the per-weight linear map
a -> a*w is a 3x3 matrix over F_p whose entries are precomputed (7 distinct words per weight).
Weights and challenges live in constant memory (warp-uniform operands).  Chunks are packed into a
few translation units so `make -j` compiles them in parallel.

Run:  python -m airgen.codegen_cuda   (writes csrc/air_gen/*)
"""
import os
import sys

from .build import CATEGORIES, build_air
from .circuit import P, reachable_postorder

R = (1 << 64) % P
OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "csrc", "air_gen")
CHUNK_COST_BUDGET = float(os.environ.get("TVM_AIR_BUDGET", "100"))
SYNC_EVERY = int(os.environ.get("TVM_AIR_SYNC_EVERY", "0"))
NUM_TUS = int(os.environ.get("TVM_AIR_TUS", "8"))
MIN_BLOCKS = int(os.environ.get("TVM_AIR_MIN_BLOCKS", "1"))
WTAB_WORDS = 7   # per weight: b0, b1, b2, -b1, -b2, b0+b2, b1-b2


def mont(v):
    return v * R % P


def node_cost(b, n):
    """rough instruction weight of one operation (B-field multiply = 1)"""
    if n.kind not in "+*":
        return 0.0
    lx = not b.evaluates_to_base_element(n.lhs)
    rx = not b.evaluates_to_base_element(n.rhs)
    if n.kind == "*":
        return 9.0 if (lx and rx) else (3.0 if (lx or rx) else 1.0)
    return 1.0 if (lx and rx) else 0.35


def chunk_constraints(air, budget=None):
    """-> list of (category, [(global_weight_index, node)]).  Chunks are contiguous ranges of the
    evaluator-order constraint list; the budget is in units of B-field multiplications because
    ptxas time grows super-linearly with the size of the basic block."""
    budget = budget or CHUNK_COST_BUDGET
    chunks, offset = [], 0
    for cat in CATEGORIES:
        b = air.builders[cat]
        cs = air.constraints[cat]
        cur, seen, cost = [], set(), 0.0
        for j, c in enumerate(cs):
            new = [x for x in reachable_postorder([c]) if id(x) not in seen and x.kind in "+*"]
            new_cost = sum(node_cost(b, x) for x in new) + 6.0
            if cur and cost + new_cost > budget:
                chunks.append((cat, cur))
                cur, seen, cost = [], set(), 0.0
                new = [x for x in reachable_postorder([c]) if x.kind in "+*"]
                new_cost = sum(node_cost(b, x) for x in new) + 6.0
            cur.append((offset + j, c))
            seen.update(id(x) for x in new)
            cost += new_cost
        if cur:
            chunks.append((cat, cur))
        offset += len(cs)
    return chunks


class Emitter:
    def __init__(self, builder):
        self.b = builder
        self.lines = []
        self.name = {}       # id(node) -> C expression (variable name or literal)
        self.isx = {}        # id(node) -> bool (X-field valued)
        self.n = 0

    def tmp(self):
        self.n += 1
        return f"t{self.n}"

    def is_neg_one(self, n):
        return n.kind == "B" and n.val == P - 1

    def emit_node(self, n):
        k = n.kind
        key = id(n)
        if k == "B":
            self.name[key] = f"0x{mont(n.val):016x}ULL"; self.isx[key] = False
        elif k == "X":
            v = self.tmp()
            self.lines.append(f"const xfe {v} = xmake(0x{mont(n.val[0]):016x}ULL, 0x{mont(n.val[1]):016x}ULL, 0x{mont(n.val[2]):016x}ULL);")
            self.name[key] = v; self.isx[key] = True
        elif k == "C":
            v = self.tmp()
            i = n.val
            self.lines.append(f"const xfe {v} = xmake(c_ch[{3 * i}], c_ch[{3 * i + 1}], c_ch[{3 * i + 2}]);")
            self.name[key] = v; self.isx[key] = True
        elif k == "I":
            row, is_main, col = n.val
            v = self.tmp()
            if is_main:
                src = "mn" if row else "mc"
                self.lines.append(f"const u64 {v} = {src}[(size_t){col} * a.main_stride];")
                self.isx[key] = False
            else:
                src = "an" if row else "ac"
                self.lines.append(f"const xfe {v} = xmake({src}[(size_t){3 * col} * a.aux_stride], "
                                  f"{src}[(size_t){3 * col + 1} * a.aux_stride], {src}[(size_t){3 * col + 2} * a.aux_stride]);")
                self.isx[key] = True
            self.name[key] = v
        else:
            l, r = n.lhs, n.rhs
            lx, rx = self.isx[id(l)], self.isx[id(r)]
            ln, rn = self.name[id(l)], self.name[id(r)]
            v = self.tmp()
            if k == "*":
                if self.is_neg_one(l) or self.is_neg_one(r):
                    o, ox = (r, rx) if self.is_neg_one(l) else (l, lx)
                    on = self.name[id(o)]
                    expr = f"xneg({on})" if ox else f"fneg({on})"
                    resx = ox
                elif lx and rx: expr, resx = f"xmul({ln}, {rn})", True
                elif lx: expr, resx = f"xmulb({ln}, {rn})", True
                elif rx: expr, resx = f"xmulb({rn}, {ln})", True
                else: expr, resx = f"fmul({ln}, {rn})", False
            else:
                if lx and rx: expr, resx = f"xadd({ln}, {rn})", True
                elif lx: expr, resx = f"xaddb({ln}, {rn})", True
                elif rx: expr, resx = f"xaddb({rn}, {ln})", True
                else: expr, resx = f"fadd({ln}, {rn})", False
            self.lines.append(f"const {'xfe' if resx else 'u64'} {v} = {expr};")
            self.name[key] = v; self.isx[key] = resx


TU_HEADER = """// GENERATED by triton-vm_b200/airgen/codegen_cuda.py from the re-derived AIR - do not edit.
// Replaces the build-time generated Evaluable::evaluate_{cat}_constraints of the reference
// (triton-constraint-builder/src/codegen.rs:59-269) fused with the weighted sum and zerofier
// division of all_quotients_combined (triton-vm/src/table/master_table.rs:1264-1363).
#include "../air.cuh"

namespace tvm {
namespace {
__constant__ u64 c_w[%(nw)d];   // weight table of ALL constraints: AIR_WTAB_WORDS words each (see air.cuh)
__constant__ u64 c_ch[189];     // 63 challenges (X-field, Montgomery)
"""

KERNEL_HEADER = """
__global__ void __launch_bounds__(AIR_THREADS, %(minb)d) %(kname)s(AirArgs a) {
  size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
%(guard)s
  const size_t n = (size_t)1 << a.log_n;
  const size_t coset = m >> a.log_n, k = m & (n - 1);
  const size_t m_next = (coset << a.log_n) | ((k + 1) & (n - 1));
  const u64 *mc = a.main + m, *mn = a.main + m_next;
  const u64 *ac = a.aux + m, *an = a.aux + m_next;
  (void)mn; (void)an; (void)ac; (void)mc; (void)coset;
"""

TU_FOOTER = """
void %(tuname)s_launch(const AirArgs &a, const u64 *d_wtab, const u64 *d_challenges, cudaStream_t s, unsigned long long *launches) {
  cudaMemcpyToSymbolAsync(c_w, d_wtab, sizeof(u64) * %(nw)d, 0, cudaMemcpyDeviceToDevice, s);
  cudaMemcpyToSymbolAsync(c_ch, d_challenges, sizeof(u64) * 189, 0, cudaMemcpyDeviceToDevice, s);
  unsigned grid = (unsigned)((a.nrows + AIR_THREADS - 1) / AIR_THREADS);
%(launches)s
  *launches += %(nk)d;
}
}  // namespace tvm
"""


def emit_chunk(air, idx, cat, items):
    b = air.builders[cat]
    em = Emitter(b)
    roots = [c for _, c in items]
    for n in reachable_postorder(roots):
        em.emit_node(n)
    kname = f"air_chunk_{idx:03d}_{cat}"
    if SYNC_EVERY:
        guard = "  const bool active = m < a.nrows;\n  if (!active) m = 0;   // keep every thread alive for the block-wide barriers below"
    else:
        guard = "  if (m >= a.nrows) return;\n  const bool active = true;"
    src = [KERNEL_HEADER % {"kname": kname, "guard": guard, "minb": MIN_BLOCKS}]
    body = []
    for k, line in enumerate(em.lines):
        body.append(line)
        if SYNC_EVERY and (k + 1) % SYNC_EVERY == 0:
            body.append("__syncthreads();   // instruction-fetch locality: the CTA's warps share one I-cache window")
    src.append("  " + "\n  ".join(body))
    src.append("  AirAcc acc; air_acc_zero(acc);")
    for j, c in items:
        nm = em.name[id(c)]
        if em.isx[id(c)]:
            src.append(f"  air_acc_x(acc, c_w + {WTAB_WORDS * j}, {nm});")
        else:
            src.append(f"  air_acc_b(acc, c_w + {WTAB_WORDS * j}, {nm});")
    src.append(f"  if (active) air_accumulate_{cat}(a, m, coset, air_acc_reduce(acc));")
    src.append("}")
    nops = sum(1 for n in reachable_postorder(roots) if n.kind in "+*")
    return kname, "\n".join(src), nops


def main():
    air = build_air()
    chunks = chunk_constraints(air)
    total_constraints = sum(len(air.constraints[c]) for c in CATEGORIES)
    os.makedirs(OUT_DIR, exist_ok=True)
    for f in os.listdir(OUT_DIR):
        os.remove(os.path.join(OUT_DIR, f))
    # pack chunks into translation units, balancing the estimated cost (compile time)
    tus = [[] for _ in range(NUM_TUS)]
    load = [0.0] * NUM_TUS
    names, total_ops = [], 0
    order = sorted(range(len(chunks)), key=lambda i: -len(chunks[i][1]))
    emitted = {}
    for idx, (cat, items) in enumerate(chunks):
        emitted[idx] = emit_chunk(air, idx, cat, items)
    for idx in order:
        t = min(range(NUM_TUS), key=lambda i: load[i])
        tus[t].append(idx)
        load[t] += emitted[idx][2] + 20
    nw = WTAB_WORDS * total_constraints
    tu_names = []
    for t, idxs in enumerate(tus):
        if not idxs:
            continue
        idxs.sort()
        tuname = f"air_tu_{t:02d}"
        tu_names.append(tuname)
        parts = [TU_HEADER % {"nw": nw}, "}  // namespace"]
        launches = []
        for idx in idxs:
            kname, src, nops = emitted[idx]
            cat, items = chunks[idx]
            names.append((kname, cat, len(items), nops))
            total_ops += nops
            parts.append(src)
            launches.append(f"  {kname}<<<grid, AIR_THREADS, 0, s>>>(a);")
        parts.append(TU_FOOTER % {"tuname": tuname, "nw": nw, "launches": "\n".join(launches), "nk": len(idxs)})
        with open(os.path.join(OUT_DIR, tuname + ".cu"), "w") as f:
            f.write("\n".join(parts))
    unique_ops = sum(sum(1 for n in reachable_postorder(air.constraints[c]) if n.kind in "+*") for c in CATEGORIES)
    with open(os.path.join(OUT_DIR, "air_chunks.inc"), "w") as f:
        f.write("// GENERATED by airgen/codegen_cuda.py - do not edit.\n")
        f.write(f"// {len(names)} chunk kernels in {len(tu_names)} translation units, {total_ops} binary operations emitted "
                f"({unique_ops} unique in the circuit), chunk budget {CHUNK_COST_BUDGET:g}\n")
        for tuname in tu_names:
            f.write(f"TVM_AIR_TU({tuname})\n")
        for kname, cat, ncons, nops in sorted(names):
            f.write(f"// {kname}: {ncons} constraints, {nops} ops\n")
    with open(os.path.join(OUT_DIR, "air_meta.inc"), "w") as f:
        f.write("// GENERATED by airgen/codegen_cuda.py - constraint degrees in evaluator order\n")
        f.write("// (the generated *_quotient_degree_bounds of the reference, codegen.rs:222-229).\n")
        for cat in CATEGORIES:
            b = air.builders[cat]
            degs = [b.degree(c) for c in air.constraints[cat]]
            f.write(f"static const int AIR_NUM_{cat.upper()} = {len(degs)};\n")
            f.write(f"static const unsigned char AIR_DEGREES_{cat.upper()}[] = {{{', '.join(map(str, degs))}}};\n")
    print(f"wrote {len(names)} chunks in {len(tu_names)} TUs to {OUT_DIR}: {total_ops} ops emitted, {unique_ops} unique")


if __name__ == "__main__":
    main()

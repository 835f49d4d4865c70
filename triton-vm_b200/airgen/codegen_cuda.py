"""Emit sm_100a CUDA source for the AIR quotient kernels from the lowered constraint circuits.

Reference behaviour being replaced: `all_quotients_combined` (triton-vm/src/table/
master_table.rs:1264-1363) calling the build-time generated `Evaluable::evaluate_*_constraints`
(triton-constraint-builder/src/codegen.rs:59-269).  For every quotient-domain row the result is
    sum_cat  zerofier_inverse_cat(row) * sum_j  weight_j * constraint_j(row, next_row)
with constraint order init | cons | tran | term and, inside a category, base-field-valued
constraints before extension-field-valued ones (codegen.rs:210-212).

B200 mapping: one thread per row, straight-line code with common sub-expressions shared inside a
chunk.  The 404 transition constraints are split into several kernels ("chunks") so that each is a
basic block ptxas compiles in seconds; a node needed by two chunks is recomputed.  Every chunk
adds its zerofier-weighted partial sum into the row-planar X-field output.  Weights and
challenges live in constant memory (warp-uniform operands).

Run:  python -m airgen.codegen_cuda   (writes csrc/air_gen/*)
"""
import os
import sys

from .build import CATEGORIES, build_air
from .circuit import P, reachable_postorder

R = (1 << 64) % P
OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "csrc", "air_gen")
CHUNK_COST_BUDGET = 1300.0
SYNC_EVERY = int(os.environ.get("TVM_AIR_SYNC_EVERY", "40"))


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


HEADER = """// GENERATED by triton-vm_b200/airgen/codegen_cuda.py from the re-derived AIR — do not edit.
// Replaces the build-time generated Evaluable::evaluate_{cat}_constraints of the reference
// (triton-constraint-builder/src/codegen.rs:59-269) fused with the weighted sum and zerofier
// division of all_quotients_combined (triton-vm/src/table/master_table.rs:1264-1363).
#include "../air.cuh"

namespace tvm {
namespace {
__constant__ u64 c_w[%(nw)d];   // this chunk's quotient weights (X-field, Montgomery), 3 words each
__constant__ u64 c_ch[189];  // 63 challenges (X-field, Montgomery)
}

__global__ void __launch_bounds__(AIR_THREADS) %(kname)s(AirArgs a) {
  size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = m < a.nrows;
  if (!active) m = 0;              // keep every thread alive for the block-wide barriers below
  const size_t n = (size_t)1 << a.log_n;
  const size_t coset = m >> a.log_n, k = m & (n - 1);
  const size_t m_next = (coset << a.log_n) | ((k + 1) & (n - 1));
  const u64 *mc = a.main + m, *mn = a.main + m_next;
  const u64 *ac = a.aux + m, *an = a.aux + m_next;
  (void)mn; (void)an; (void)ac; (void)mc;
"""

FOOTER = """
void %(kname)s_launch(const AirArgs &a, const u64 *d_weights_all, const u64 *d_challenges, cudaStream_t s) {
  // this chunk covers the contiguous constraint range [%(wstart)d, %(wstart)d + %(wcount)d) of the weight vector
  cudaMemcpyToSymbolAsync(c_w, d_weights_all + 3 * %(wstart)d, sizeof(u64) * 3 * %(wcount)d, 0, cudaMemcpyDeviceToDevice, s);
  cudaMemcpyToSymbolAsync(c_ch, d_challenges, sizeof(u64) * 189, 0, cudaMemcpyDeviceToDevice, s);
  unsigned grid = (unsigned)((a.nrows + AIR_THREADS - 1) / AIR_THREADS);
  %(kname)s<<<grid, AIR_THREADS, 0, s>>>(a);
}
}  // namespace tvm
"""


def emit_chunk(air, idx, cat, items):
    b = air.builders[cat]
    em = Emitter(b)
    roots = [c for _, c in items]
    for n in reachable_postorder(roots):
        em.emit_node(n)
    kname = f"air_chunk_{idx:02d}_{cat}"
    src = [HEADER % {"nw": 3 * len(items), "kname": kname}]
    body = []
    for k, line in enumerate(em.lines):
        body.append(line)
        if SYNC_EVERY and (k + 1) % SYNC_EVERY == 0:
            body.append("__syncthreads();   // instruction-fetch locality: the CTA's warps share one I-cache window")
    src.append("  " + "\n  ".join(body))
    src.append("  xfe acc = xzero();")
    for slot, (_, c) in enumerate(items):
        w = f"xmake(c_w[{3 * slot}], c_w[{3 * slot + 1}], c_w[{3 * slot + 2}])"
        nm = em.name[id(c)]
        if em.isx[id(c)]:
            src.append(f"  acc = xadd(acc, xmul({w}, {nm}));")
        else:
            src.append(f"  acc = xadd(acc, xmulb({w}, {nm}));")
    src.append(f"  if (active) air_accumulate_{cat}(a, m, coset, k, acc);")
    src.append("}")
    idxs = [j for j, _ in items]
    assert idxs == list(range(idxs[0], idxs[0] + len(idxs)))
    src.append(FOOTER % {"kname": kname, "wstart": idxs[0], "wcount": len(idxs)})
    nops = sum(1 for n in reachable_postorder(roots) if n.kind in "+*")
    return kname, "\n".join(src), nops


def main():
    air = build_air()
    chunks = chunk_constraints(air)
    os.makedirs(OUT_DIR, exist_ok=True)
    for f in os.listdir(OUT_DIR):
        os.remove(os.path.join(OUT_DIR, f))
    names, total_ops = [], 0
    for idx, (cat, items) in enumerate(chunks):
        kname, src, nops = emit_chunk(air, idx, cat, items)
        names.append((kname, cat, len(items), nops))
        total_ops += nops
        with open(os.path.join(OUT_DIR, kname + ".cu"), "w") as f:
            f.write(src)
    unique_ops = sum(sum(1 for n in reachable_postorder(air.constraints[c]) if n.kind in "+*") for c in CATEGORIES)
    with open(os.path.join(OUT_DIR, "air_chunks.inc"), "w") as f:
        f.write("// GENERATED by airgen/codegen_cuda.py — do not edit.\n")
        f.write(f"// {len(names)} chunks, {total_ops} binary operations emitted ({unique_ops} unique in the circuit)\n")
        for kname, cat, ncons, nops in names:
            f.write(f"TVM_AIR_CHUNK({kname}) // {cat}: {ncons} constraints, {nops} ops\n")
    with open(os.path.join(OUT_DIR, "air_meta.inc"), "w") as f:
        f.write("// GENERATED by airgen/codegen_cuda.py — constraint degrees in evaluator order\n")
        f.write("// (the generated *_quotient_degree_bounds of the reference, codegen.rs:222-229).\n")
        for cat in CATEGORIES:
            b = air.builders[cat]
            degs = [b.degree(c) for c in air.constraints[cat]]
            f.write(f"static const int AIR_NUM_{cat.upper()} = {len(degs)};\n")
            f.write(f"static const unsigned char AIR_DEGREES_{cat.upper()}[] = {{{', '.join(map(str, degs))}}};\n")
    print(f"wrote {len(names)} chunks to {OUT_DIR}: {total_ops} ops emitted, {unique_ops} unique")
    for n in names:
        print("  ", n)


if __name__ == "__main__":
    main()

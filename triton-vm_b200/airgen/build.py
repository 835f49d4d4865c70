"""Constraints::all() + degree lowering — restates triton-constraint-builder/src/lib.rs:38-184
and the build script triton-vm/build.rs:11-25.

`build_air()` returns, per category (init, cons, tran, term):
  * the final constraint list in evaluator order: originals ++ main substitutions ++ aux
    substitutions (lib.rs:174-184), then stably partitioned base-field-valued first
    (codegen.rs:210-212, 243-252);
  * the substitution rules, from which the derived (degree-lowering) columns are filled
    (substitutions.rs:128-301).
"""
import time

from . import cross_table
from .circuit import Builder, lower_to_degree, num_visible_nodes, multicircuit_degree
from .columns import NUM_MAIN_COLUMNS, NUM_AUX_COLUMNS
from .tables import program, processor, op_stack, ram, jump_stack, hash, cascade, lookup, u32

TARGET_DEGREE = 4            # triton-air/src/lib.rs:37
PROVIDERS = [program, processor, op_stack, ram, jump_stack, hash, cascade, lookup, u32, cross_table]
CATEGORIES = ["init", "cons", "tran", "term"]
_FN = {"init": "initial", "cons": "consistency", "tran": "transition", "term": "terminal"}

# specification/src/arithmetization-overview.md:28-78
EXPECTED_BEFORE = {"init": (79, 539), "cons": (79, 637), "tran": (152, 6825), "term": (23, 213)}
EXPECTED_AFTER = {"init": (81, 543), "cons": (97, 689), "tran": (403, 7400), "term": (23, 213)}


class Air:
    pass


def build_air(verbose=False):
    air = Air()
    air.builders, air.roots, air.before = {}, {}, {}
    for cat in CATEGORIES:
        b = Builder(dual=(cat == "tran"))
        roots = []
        for prov in PROVIDERS:
            roots += [m.n for m in getattr(prov, _FN[cat])(b)]
        air.builders[cat], air.roots[cat] = b, roots
        air.before[cat] = (len(roots), num_visible_nodes(roots), multicircuit_degree(b, roots))
        if verbose:
            print(cat, "before lowering: constraints, nodes, degree =", air.before[cat], flush=True)
    # lower_to_target_degree_through_substitutions (lib.rs:131-171): column counters carry over
    n_main, n_aux = NUM_MAIN_COLUMNS, NUM_AUX_COLUMNS
    air.main_subst, air.aux_subst, air.after, air.constraints = {}, {}, {}, {}
    air.subst_col_start = {}
    for cat in CATEGORIES:
        t0 = time.time()
        b, roots = air.builders[cat], air.roots[cat]
        air.subst_col_start[cat] = (n_main, n_aux)
        ms, xs = lower_to_degree(b, roots, TARGET_DEGREE, n_main, n_aux)
        n_main += len(ms); n_aux += len(xs)
        air.main_subst[cat], air.aux_subst[cat] = ms, xs
        allc = roots + ms + xs
        air.after[cat] = (len(allc), num_visible_nodes(allc), multicircuit_degree(b, allc))
        base = [c for c in allc if b.evaluates_to_base_element(c)]
        ext = [c for c in allc if not b.evaluates_to_base_element(c)]
        air.constraints[cat] = base + ext
        if verbose:
            print(cat, "after lowering:", air.after[cat], f"(+{len(ms)} main, +{len(xs)} aux columns)",
                  f"{time.time() - t0:.1f}s", flush=True)
    air.num_main_columns, air.num_aux_columns = n_main, n_aux
    return air


if __name__ == "__main__":
    a = build_air(verbose=True)
    print("main columns", a.num_main_columns, "aux columns", a.num_aux_columns)
    for cat in CATEGORIES:
        print(cat, "before ok" if a.before[cat][:2] == EXPECTED_BEFORE[cat] else f"BEFORE MISMATCH {EXPECTED_BEFORE[cat]}",
              "| after ok" if a.after[cat][:2] == EXPECTED_AFTER[cat] else f"| AFTER MISMATCH {EXPECTED_AFTER[cat]}")

"""CPU tests of the re-derived AIR (triton-vm_b200/airgen): constraint / node / column counts
against the reference's specification, and the reference's own golden fingerprint of the
generated evaluators (`air_constraints_evaluators_have_not_changed`, master_table.rs:2327-2415)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "triton-vm_b200"))

from airgen.build import CATEGORIES, EXPECTED_BEFORE, PROVIDERS, _FN, build_air  # noqa: E402
from airgen.circuit import Builder, lower_to_degree, reachable_postorder, xadd, xmul  # noqa: E402
from airgen.columns import AUX, MAIN  # noqa: E402
from airgen.evaluate import evaluate_constraints  # noqa: E402
from oracle.rand_compat import StdRng  # noqa: E402


@pytest.fixture(scope="module")
def air():
    return build_air()


def test_counts_before_lowering_match_specification(air):
    # specification/src/arithmetization-overview.md:28-47: 79/79/152/23 constraints, 539/637/6825/213 nodes
    for cat in CATEGORIES:
        assert air.before[cat][:2] == EXPECTED_BEFORE[cat]
    assert [air.before[c][2] for c in CATEGORIES] == [8, 11, 19, 4]


def test_counts_after_lowering(air):
    # arithmetization-overview.md:62-78 and :7-20: 81/97/403/23 constraints; 230 + 41 derived columns
    assert [air.after[c][0] for c in CATEGORIES] == [81, 97, 403, 23]
    assert all(air.after[c][2] <= 4 for c in CATEGORIES)
    assert air.num_main_columns == 379 and air.num_aux_columns == 90  # +1 batch-randomizer column = 91
    assert sum(len(air.main_subst[c]) for c in CATEGORIES) == 230
    assert sum(len(air.aux_subst[c]) for c in CATEGORIES) == 41


def _global_key(intern, n, memo):
    stack = [n]
    while stack:
        x = stack[-1]
        if id(x) in memo:
            stack.pop(); continue
        if x.kind in "+*":
            kl, kr = memo.get(id(x.lhs)), memo.get(id(x.rhs))
            if kl is None: stack.append(x.lhs)
            if kr is None: stack.append(x.rhs)
            if kl is None or kr is None: continue
            raw = (x.kind, kl, kr)
        else:
            raw = (x.kind, x.val)
        memo[id(x)] = intern.setdefault(raw, len(intern))
        stack.pop()
    return memo[id(n)]


@pytest.mark.parametrize("target,expected", [
    (8, {"init": (79, 539), "cons": (83, 648), "tran": (263, 7059), "term": (23, 213)}),
    (4, {"init": (81, 543), "cons": (97, 689), "tran": (403, 7400), "term": (23, 213)}),
])
def test_per_table_lowering_reproduces_specification_node_counts(target, expected):
    """The specification's post-lowering table (arithmetization-overview.md:49-78) is produced by
    lowering every table separately with its own column offsets (master_table.rs:1925-2030); the
    node counts are sensitive to which nodes the lowering heuristic picks."""
    names = ["program", "processor", "op_stack", "ram", "jump_stack", "hash", "cascade", "lookup", "u32"]
    ends = [(MAIN[n].start + MAIN[n].COUNT, AUX[n].start + AUX[n].COUNT) for n in names] + [(0, 0)]
    for cat in CATEGORIES:
        intern, keys, total = {}, set(), 0
        for prov, (me, ae) in zip(PROVIDERS, ends):
            b = Builder(dual=(cat == "tran"))
            roots = [m.n for m in getattr(prov, _FN[cat])(b)]
            orig = list(roots)
            ms, xs = lower_to_degree(b, roots, target, me, ae)
            allc = orig + ms + xs
            total += len(allc)
            memo = {}
            for x in reachable_postorder(allc):
                keys.add(_global_key(intern, x, memo))
        assert (total, len(keys)) == expected[cat], cat


def test_air_evaluator_fingerprint_matches_reference_golden_value(air):
    # master_table.rs:2327-2415: seeded random rows/challenges, all 8 evaluators, one XFE fingerprint
    rng = StdRng.seed_from_u64(3508729174085202315)
    nm, na = 379, 91
    mcb = [rng.bfe() for _ in range(nm)]; mce = [rng.xfe() for _ in range(nm)]; ac = [rng.xfe() for _ in range(na)]
    mnb = [rng.bfe() for _ in range(nm)]; mne = [rng.xfe() for _ in range(nm)]; an = [rng.xfe() for _ in range(na)]
    ch = [rng.xfe() for _ in range(63)]
    coeffs = []
    for cat in CATEGORIES:
        cs = air.constraints[cat]
        coeffs += evaluate_constraints(cs, mcb, ac, mnb, an, ch)
        coeffs += evaluate_constraints(cs, mce, ac, mne, an, ch)
    x = rng.xfe()
    acc = (0, 0, 0)
    for c in reversed(coeffs):
        acc = xadd(xmul(acc, x), c)
    assert acc == (17974882881108171077, 15638927082579294872, 9717283721935042729)

"""Node-by-node evaluation of the constraint circuits on Python ints (canonical field elements).
Used by the tests as the reference evaluator (mirrors the semantics of the generated
`Evaluable::evaluate_*_constraints`, triton-constraint-builder/src/codegen.rs:59-269:
base-field-valued constraints first, then extension-field-valued ones)."""
from .circuit import P, xmul, xadd, reachable_postorder


def _lift(v):
    return v if isinstance(v, tuple) else (v % P, 0, 0)


def evaluate_constraints(constraints, cur_main, cur_aux, next_main, next_aux, challenges):
    """Rows are lists of ints (base) or 3-tuples (extension); returns a list of 3-tuples."""
    val = {}
    for n in reachable_postorder(constraints):
        k = n.kind
        if k == "B": v = (n.val, 0, 0)
        elif k == "X": v = n.val
        elif k == "C": v = _lift(challenges[n.val])
        elif k == "I":
            row, is_main, col = n.val
            src = (next_main if row else cur_main) if is_main else (next_aux if row else cur_aux)
            v = _lift(src[col])
        elif k == "+": v = xadd(val[id(n.lhs)], val[id(n.rhs)])
        else: v = xmul(val[id(n.lhs)], val[id(n.rhs)])
        val[id(n)] = v
    return [val[id(c)] for c in constraints]

"""Constraint-circuit builder with the exact node-identity semantics of
triton-constraint-circuit/src/lib.rs (reference @ 8cd9a0eb):

* leaves BConst / XConst (demoted to BConst when unlift() succeeds, lib.rs:1087-1092) /
  Input / Challenge; inner nodes Add / Mul only (lib.rs:293-299);
* `a - b` = `a + (-1)*b`, `-a` = `(-1)*a` (lib.rs:742-764); Sum folds left (768-773);
* binop(): neutral-element shortcuts, constant folding, structural dedup trying the commuted
  operand order first, otherwise a fresh node whose id is the builder's monotone counter
  (lib.rs:666-720); leaves dedup the same way (1094-1106);
* degree() (505-530), evaluates_to_base_element() (585-595);
* lower_to_degree / pick_node_to_substitute / apply_substitution /
  redirect_all_references_to_node (820-958, 1114-1129).

Node ids matter: degree lowering breaks ties by smallest id (lib.rs:936-957), so every leaf and
operator must be created in the same order as in the reference's table definitions.
"""
P = (1 << 64) - (1 << 32) + 1


def xmul(a, b):
    d0 = a[0] * b[0]; d1 = a[0] * b[1] + a[1] * b[0]; d2 = a[0] * b[2] + a[1] * b[1] + a[2] * b[0]
    d3 = a[1] * b[2] + a[2] * b[1]; d4 = a[2] * b[2]
    return ((d0 - d3) % P, (d1 + d3 - d4) % P, (d2 + d4) % P)


def xadd(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P, (a[2] + b[2]) % P)


class Node:
    __slots__ = ("id", "kind", "val", "lhs", "rhs")
    # kind: 'B' (val=int), 'X' (val=(c0,c1,c2)), 'I' (val=(row, is_main, col)), 'C' (val=int),
    #       '+' / '*' (lhs, rhs)

    def __init__(self, id_, kind, val=None, lhs=None, rhs=None):
        self.id, self.kind, self.val, self.lhs, self.rhs = id_, kind, val, lhs, rhs

    def is_binop(self):
        return self.kind in "+*"

    def is_zero(self):
        return (self.kind == "B" and self.val == 0) or (self.kind == "X" and self.val == (0, 0, 0))

    def is_one(self):
        return (self.kind == "B" and self.val == 1) or (self.kind == "X" and self.val == (1, 0, 0))


class Builder:
    def __init__(self, dual):
        self.dual = dual              # DualRowIndicator (transition) vs SingleRowIndicator
        self.id_counter = 0
        self.all_nodes = {}           # id -> Node
        self._intern = {}             # structural key interning
        self._skey_memo = {}          # id(node) -> structural key (cleared on substitution)
        self._by_skey = {}            # structural key -> Node, for nodes in all_nodes
        self._deg_memo = {}
        self._base_memo = {}

    # -- structural equality (lib.rs:331-344, 372-380) ---------------------------------
    def skey(self, n):
        k = self._skey_memo.get(id(n))
        if k is not None:
            return k
        # iterative post-order to avoid deep recursion
        stack = [n]
        memo = self._skey_memo
        while stack:
            x = stack[-1]
            if id(x) in memo:
                stack.pop(); continue
            if x.kind in "+*":
                l, r = x.lhs, x.rhs
                kl, kr = memo.get(id(l)), memo.get(id(r))
                if kl is None: stack.append(l)
                if kr is None: stack.append(r)
                if kl is None or kr is None:
                    continue
                raw = (x.kind, kl, kr)
            else:
                raw = (x.kind, x.val)
            k = self._intern.get(raw)
            if k is None:
                k = len(self._intern); self._intern[raw] = k
            memo[id(x)] = k
            stack.pop()
        return memo[id(n)]

    def _rebuild_index(self):
        self._skey_memo = {}
        self._deg_memo = {}
        self._base_memo = {}
        self._by_skey = {}
        for n in self.all_nodes.values():
            self._by_skey.setdefault(self.skey(n), n)

    def _find_equal(self, kind, val=None, lhs=None, rhs=None):
        if kind in "+*":
            raw = (kind, self.skey(lhs), self.skey(rhs))
        else:
            raw = (kind, val)
        k = self._intern.get(raw)
        if k is None:
            return None
        return self._by_skey.get(k)

    def _insert(self, node):
        assert node.id not in self.all_nodes
        self.all_nodes[node.id] = node
        self._by_skey[self.skey(node)] = node
        self.id_counter += 1
        return node

    # -- leaves (lib.rs:1048-1112) -------------------------------------------------------
    def _make_leaf(self, kind, val):
        if kind == "X" and val[1] == 0 and val[2] == 0:
            kind, val = "B", val[0]
        same = self._find_equal(kind, val)
        if same is not None:
            return same
        return self._insert(Node(self.id_counter, kind, val))

    def b_constant(self, v):
        return M(self, self._make_leaf("B", v % P))

    def x_constant(self, v):
        if isinstance(v, int):
            v = (v % P, 0, 0)
        return M(self, self._make_leaf("X", tuple(c % P for c in v)))

    def zero(self): return self.b_constant(0)
    def one(self): return self.b_constant(1)
    def minus_one(self): return self.b_constant(P - 1)

    def input(self, row, is_main, col):
        if not self.dual:
            assert row == 0
        return M(self, self._make_leaf("I", (row, bool(is_main), col)))

    def challenge(self, idx):
        return M(self, self._make_leaf("C", int(idx)))

    # -- binop (lib.rs:666-720) ------------------------------------------------------------
    def binop(self, op, lhs, rhs):
        if op == "+":
            if rhs.is_zero(): return lhs
            if lhs.is_zero(): return rhs
        else:
            if rhs.is_one(): return lhs
            if lhs.is_one(): return rhs
            if rhs.is_zero(): return rhs
            if lhs.is_zero(): return lhs
        if lhs.kind in "BX" and rhs.kind in "BX":
            if lhs.kind == "B" and rhs.kind == "B":
                v = (lhs.val + rhs.val) % P if op == "+" else (lhs.val * rhs.val) % P
                return self._make_leaf("B", v)
            l = lhs.val if lhs.kind == "X" else (lhs.val, 0, 0)
            r = rhs.val if rhs.kind == "X" else (rhs.val, 0, 0)
            return self._make_leaf("X", xadd(l, r) if op == "+" else xmul(l, r))
        n = self._find_equal(op, lhs=rhs, rhs=lhs)
        if n is not None: return n
        n = self._find_equal(op, lhs=lhs, rhs=rhs)
        if n is not None: return n
        return self._insert(Node(self.id_counter, op, None, lhs, rhs))

    # -- analysis --------------------------------------------------------------------------
    def degree(self, n):
        memo = self._deg_memo
        d = memo.get(id(n))
        if d is not None: return d
        stack = [n]
        while stack:
            x = stack[-1]
            if id(x) in memo:
                stack.pop(); continue
            if x.kind in "+*":
                dl, dr = memo.get(id(x.lhs)), memo.get(id(x.rhs))
                if dl is None: stack.append(x.lhs)
                if dr is None: stack.append(x.rhs)
                if dl is None or dr is None: continue
                if x.kind == "+": d = max(dl, dr)
                else: d = -1 if min(dl, dr) <= -1 else dl + dr
            elif x.is_zero(): d = -1
            elif x.kind == "I": d = 1
            else: d = 0
            memo[id(x)] = d
            stack.pop()
        return memo[id(n)]

    def evaluates_to_base_element(self, n):
        memo = self._base_memo
        b = memo.get(id(n))
        if b is not None: return b
        stack = [n]
        while stack:
            x = stack[-1]
            if id(x) in memo:
                stack.pop(); continue
            if x.kind in "+*":
                bl, br = memo.get(id(x.lhs)), memo.get(id(x.rhs))
                if bl is None: stack.append(x.lhs)
                if br is None: stack.append(x.rhs)
                if bl is None or br is None: continue
                b = bl and br
            elif x.kind == "B": b = True
            elif x.kind == "I": b = x.val[1]
            else: b = False
            memo[id(x)] = b
            stack.pop()
        return memo[id(n)]


class M:
    """ConstraintCircuitMonad: node + builder, with the operator overloads of lib.rs:722-774."""
    __slots__ = ("b", "n")

    def __init__(self, b, n):
        self.b, self.n = b, n

    def _c(self, o):
        if isinstance(o, M):
            assert o.b is self.b
            return o
        raise TypeError(o)

    def __add__(self, o): return M(self.b, self.b.binop("+", self.n, self._c(o).n))
    def __mul__(self, o): return M(self.b, self.b.binop("*", self.n, self._c(o).n))
    def __neg__(self): return M(self.b, self.b.binop("*", self.b.minus_one().n, self.n))
    def __sub__(self, o):
        o = self._c(o)
        return M(self.b, self.b.binop("+", self.n, (-o).n))
    def clone(self): return self


def msum(items):
    """impl Sum: reduce(|a, i| a + i); panics on empty (lib.rs:766-773).  Lazy like Rust's
    iterator chain: item k+1 is produced only after the partial sum over items 0..k exists."""
    it = iter(items)
    acc = next(it)
    for item in it:
        acc = acc + item
    return acc


# ---- multicircuit utilities --------------------------------------------------------------
def reachable_postorder(roots):
    """Unique nodes reachable from roots; children before parents."""
    seen, order = set(), []
    for r in roots:
        if id(r) in seen: continue
        stack = [(r, 0)]
        while stack:
            x, st = stack.pop()
            if st == 0:
                if id(x) in seen: continue
                seen.add(id(x))
                stack.append((x, 1))
                if x.kind in "+*":
                    stack.append((x.rhs, 0)); stack.append((x.lhs, 0))
            else:
                order.append(x)
    return order


def num_visible_nodes(roots):
    return len(reachable_postorder(roots))


def multicircuit_degree(b, roots):
    return max((b.degree(r) for r in roots), default=-1)


def pick_node_to_substitute(b, roots, target_degree):
    """lib.rs:902-958.  Occurrence counts are taken over the tree expansion (with multiplicity)
    of every distinct high-degree node; computed here as DAG path counts."""
    order = reachable_postorder(roots)           # children first
    high = {id(x) for x in order if b.degree(x) > target_degree}
    f = {id(x): (1 if id(x) in high else 0) for x in order}
    for x in reversed(order):                    # parents before children
        fx = f[id(x)]
        if fx and x.kind in "+*":
            f[id(x.lhs)] += fx
            f[id(x.rhs)] += fx
    best = None
    for x in order:
        d = b.degree(x)
        if 1 < d <= target_degree and f[id(x)] > 0:
            key = (f[id(x)], d, -x.id)
            if best is None or key > best[0]:
                best = (key, x)
    assert best is not None, "Cannot lower degree."
    return best[1]


def apply_substitution(b, roots, num_main_cols, num_aux_cols, chosen, n_main_so_far, n_aux_so_far):
    """lib.rs:864-897 + 1114-1129."""
    is_main = b.evaluates_to_base_element(chosen)
    if is_main:
        new_var = b.input(0, True, num_main_cols + n_main_so_far)
    else:
        new_var = b.input(0, False, num_aux_cols + n_aux_so_far)
    # redirect_all_references_to_node
    del b.all_nodes[chosen.id]
    for n in b.all_nodes.values():
        if n.kind in "+*":
            if n.lhs.id == chosen.id: n.lhs = new_var.n
            if n.rhs.id == chosen.id: n.rhs = new_var.n
    for i, r in enumerate(roots):
        if r.id == chosen.id:
            roots[i] = new_var.n
    b._rebuild_index()
    return new_var - M(b, chosen), is_main


def lower_to_degree(b, roots, target_degree, num_main_cols, num_aux_cols):
    """lib.rs:820-856.  `roots` (list of Node) is modified in place; returns
    (main_substitution_constraints, aux_substitution_constraints) as lists of Node."""
    main_c, aux_c = [], []
    if not roots:
        return main_c, aux_c
    b._rebuild_index()
    while multicircuit_degree(b, roots) > target_degree:
        chosen = pick_node_to_substitute(b, roots, target_degree)
        new_c, _ = apply_substitution(b, roots, num_main_cols, num_aux_cols, chosen, len(main_c), len(aux_c))
        if b.evaluates_to_base_element(new_c.n):
            main_c.append(new_c.n)
        else:
            aux_c.append(new_c.n)
    return main_c, aux_c

"""Merkle tree over Tip5 digests + authentication structures (pure Python).

Restates twenty-first 2.0's `MerkleTree` as the reference uses it: `par_new`
(master_table.rs:449, stark.rs:443, fri.rs:346), `root`, `authentication_structure`
(stark.rs:676-713, fri.rs:311) and the verifier-side `MerkleTreeInclusionProof`
(stark.rs:1609-1671).  Node i = hash_pair(node 2i, node 2i+1); root = node 1; leaf j =
node n + j; the authentication structure lists the needed-but-not-computable sibling nodes in
DESCENDING node-index order (pinned by the reference's whole-proof digests, tests/test_golden.py).
TEST INFRASTRUCTURE ONLY."""
from .tip5 import hash_pair


class MerkleTree:
    def __init__(self, leaves):
        n = len(leaves)
        assert n >= 1 and n & (n - 1) == 0
        self.num_leafs = n
        self.nodes = [[0] * 5 for _ in range(2 * n)]
        for j, l in enumerate(leaves):
            self.nodes[n + j] = list(l)
        for i in range(n - 1, 0, -1):
            self.nodes[i] = hash_pair(self.nodes[2 * i], self.nodes[2 * i + 1])

    def root(self):
        return self.nodes[1]

    def authentication_structure(self, leaf_indices):
        return [self.nodes[i] for i in auth_structure_node_indices(self.num_leafs, leaf_indices)]


def auth_structure_node_indices(num_leafs, leaf_indices):
    needed, computable = set(), set()
    for li in leaf_indices:
        assert 0 <= li < num_leafs
        node = li + num_leafs
        while node > 1:
            computable.add(node)
            needed.add(node ^ 1)
            node //= 2
    return sorted(needed - computable, reverse=True)


def verify_inclusion(root, tree_height, indexed_leafs, auth_structure):
    """MerkleTreeInclusionProof::verify: recompute the root from leaves + auth structure."""
    num_leafs = 1 << tree_height
    idx = auth_structure_node_indices(num_leafs, [i for i, _ in indexed_leafs])
    if len(idx) != len(auth_structure):
        return False
    known = {}
    for i, d in zip(idx, auth_structure):
        known[i] = list(d)
    for i, leaf in indexed_leafs:
        node = i + num_leafs
        if node in known and known[node] != list(leaf):
            return False
        known[node] = list(leaf)
    level = sorted(set(i + num_leafs for i, _ in indexed_leafs), reverse=True)
    frontier = set(level)
    while frontier != {1}:
        nxt = set()
        for node in frontier:
            if node == 1:
                nxt.add(1)
                continue
            sib = node ^ 1
            if sib not in known:
                return False
            l, r = (known[node], known[sib]) if node % 2 == 0 else (known[sib], known[node])
            parent = node // 2
            h = hash_pair(l, r)
            if parent in known and known[parent] != h:
                return False
            known[parent] = h
            nxt.add(parent)
        frontier = nxt
    return known[1] == list(root)

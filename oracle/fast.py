"""Complete Prover::prove (stark.rs:331-719) with every array stage in C + OpenMP (oracle/c): the same transcript, the same
proof words as oracle/stark.py's `prove`, whose Python-int loops it replaces stage by stage (tests/test_oracle_fast.py
compares the two at small sizes, FRI and STIR).  It exists so that a COMPLETE prove of the CPU restatement can be timed
on the host cores at BASELINE heights (bench.py --impl reference; `timings` collects seconds per stage under the
reference profiler's labels, profiler.rs:10-31).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the product path never imports this.

Big arrays stay in Montgomery form (numpy uint64) between the C calls; X-field codewords are [len, 3]."""
import time

import numpy as np

from . import codec, corc, field as F, merkle, stark as S, stir as ST
from .field import P


def _can(a):
    return corc.from_mont(np.ascontiguousarray(a))


def _xlist(a_m):
    """Montgomery [k,3] -> list of canonical X-field tuples"""
    return [tuple(int(v) for v in r) for r in _can(a_m).reshape(-1, 3)]


def _digest(nodes_m, k=1):
    return [int(v) for v in _can(nodes_m[k])]


def _auth(nodes_m, num_leafs, indices):
    idx = merkle.auth_structure_node_indices(num_leafs, indices)
    if not idx:
        return []
    return [[int(v) for v in row] for row in _can(nodes_m[np.array(idx)]).reshape(-1, 5)]


def _xfe_leaves(cw_m):
    return np.concatenate([cw_m, np.zeros((cw_m.shape[0], 2), dtype=np.uint64)], axis=1)


class _Clock:
    def __init__(self, sink):
        self.sink, self.t = sink, time.perf_counter()

    def mark(self, label):
        now = time.perf_counter()
        if self.sink is not None:
            self.sink[label] = self.sink.get(label, 0.0) + (now - self.t)
        self.t = now


# ---- FRI (fri.rs:212-366, 754-772) -------------------------------------------------------------------------------
def fri_prove(ps, cw_m, d):
    rounds = []
    offset, length = d["ldt_offset"], d["ldt_len"]
    nodes = corc.merkle_build(_xfe_leaves(cw_m), mont_io=True)
    rounds.append((cw_m, nodes))
    ps.enqueue("MerkleRoot", _digest(nodes))
    for _ in range(d["fri_num_rounds"]):
        chal = ps.sample_scalars(1)[0]
        cw_m = corc.fri_fold(cw_m, offset, chal, mont_io=True)
        offset = offset * offset % P
        nodes = corc.merkle_build(_xfe_leaves(cw_m), mont_io=True)
        rounds.append((cw_m, nodes))
        ps.enqueue("MerkleRoot", _digest(nodes))
    last = rounds[-1][0]
    ps.enqueue("FriCodeword", _xlist(last))
    ps.enqueue("Polynomial", _xlist(corc.xcoset_interpolate_m(last, 1)))
    a_indices = ps.sample_indices(length, d["num_collinearity_checks"])

    def reveal(rnd, idx):
        c, nd = rounds[rnd]
        ps.enqueue("FriResponse", (_xlist(c[np.array(idx)]), _auth(nd, c.shape[0], idx)))

    reveal(0, a_indices)
    for rnd in range(len(rounds) - 1):
        n = rounds[rnd][0].shape[0]
        reveal(rnd, [(a + n // 2) % n for a in a_indices])
    ps.sample_scalars(1)
    return a_indices


# ---- STIR (stir.rs:885-993) ----------------------------------------------------------------------------------------
class _StirTree:
    """stacked Merkle tree (stir.rs:1374-1433): leaf i = hash of the ff codeword elements i, i + dist, ..."""
    def __init__(self, cw_m, ff):
        self.cw, self.ff = cw_m, ff
        self.dist = cw_m.shape[0] // ff
        table = np.ascontiguousarray(cw_m.reshape(ff, self.dist, 3).transpose(0, 2, 1)).reshape(3 * ff, self.dist)
        self.nodes = corc.merkle_build(corc.hash_rows_colmajor(table, mont_io=True), mont_io=True)

    def root(self):
        return _digest(self.nodes)

    def inclusion_proof(self, indices):
        rows = np.array([[i + j * self.dist for j in range(self.ff)] for i in indices]).reshape(-1)
        vals = _xlist(self.cw[rows])
        leafs = [vals[k * self.ff:(k + 1) * self.ff] for k in range(len(indices))]
        return leafs, _auth(self.nodes, self.dist, indices)


def stir_prove(ps, cw_m, sp):
    offset, length = sp["initial_offset"], sp["initial_domain_len"]
    assert cw_m.shape[0] == length
    ff = sp["folding_factor"]
    commitment = _StirTree(cw_m, ff)
    ps.enqueue("MerkleRoot", commitment.root())
    poly = corc.xpoly_trim_m(corc.xcoset_interpolate_m(cw_m, offset))
    first_round_indices = None
    for in_domain, ood in sp["round_queries"]:
        r = ps.sample_scalars(1)[0]
        folded = corc.xpoly_trim_m(corc.xpoly_fold_m(poly, ff, r))
        n_off, n_len = ST.next_round_domain(offset, length)
        assert folded.shape[0] <= n_len
        folded_commitment = _StirTree(corc.xcoset_evaluate_m(folded, n_off, n_len), ff)
        ps.enqueue("MerkleRoot", folded_commitment.root())
        ood_queries = ps.sample_scalars(ood)
        ood_values = [corc.xpoly_eval_m(folded, x) for x in ood_queries]
        ps.enqueue("StirOutOfDomainValues", ood_values)
        queried = ps.sample_indices(length, in_domain)
        f_off, f_len = pow(offset, ff, P), length // ff
        folded_idx = ST.unique([i % f_len for i in queried])
        ps.enqueue("StirResponse", commitment.inclusion_proof(folded_idx))
        qvals = [ST.domain_value(f_off, f_len, i) for i in folded_idx]
        points = [F.xlift(x) for x in qvals] + list(ood_queries)
        answers = [corc.xpoly_eval_m(folded, F.xlift(x)) for x in qvals] + ood_values
        pts_m, ans_m = corc._xarr(points), corc._xarr(answers)
        ans_poly = corc.xinterpolate_m(pts_m, ans_m)
        num = folded.copy() if folded.shape[0] >= ans_poly.shape[0] else np.concatenate(
            [folded, np.zeros((ans_poly.shape[0] - folded.shape[0], 3), dtype=np.uint64)])
        neg = corc._xarr([F.xneg(F.X_ONE)])[0]
        corc.xpoly_axpy_m(num, ans_poly, tuple(int(v) for v in corc.from_mont(neg)))
        quotient = corc.xpoly_div_m(num, corc.xzerofier_m(pts_m))
        dcr = ps.sample_scalars(1)[0]
        dc, acc = [], F.X_ONE
        for _ in range(len(answers) + 1):
            dc.append(acc); acc = F.xmul(acc, dcr)
        poly = corc.xpoly_trim_m(corc.xpoly_mul_m(quotient, corc._xarr(dc)))
        offset, length, commitment = n_off, n_len, folded_commitment
        if first_round_indices is None:
            first_round_indices = queried
    r = ps.sample_scalars(1)[0]
    final_poly = corc.xpoly_trim_m(corc.xpoly_fold_m(poly, ff, r))
    ps.enqueue("Polynomial", _xlist(final_poly))
    f_len = length // ff
    queried = ps.sample_indices(length, sp["final_num_in_domain_queries"])
    ps.enqueue("StirResponse", commitment.inclusion_proof(ST.unique([i % f_len for i in queried])))
    return first_round_indices if first_round_indices is not None else queried


# ---- Prover::prove ---------------------------------------------------------------------------------------------------
def prove(stark, claim, main_trace, main_rand, aux_provider, quot_rand, padded_height=None, timings=None):
    """Same contract and same proof words as oracle.stark.prove; `timings` (dict) receives seconds per stage."""
    NM, NA = S.NUM_MAIN_COLUMNS, S.NUM_AUX_COLUMNS
    clk = _Clock(timings)
    main_trace = np.ascontiguousarray(main_trace, dtype=np.uint64)
    n = main_trace.shape[1]
    d = stark.derive(padded_height or n)
    assert d["trace_len"] == n, (d["trace_len"], n)
    h, N, off = d["num_trace_randomizers"], d["ldt_len"], d["ldt_offset"]
    E = max(N, d["quotient_len"])
    es, qs = E // N, E // d["quotient_len"]
    log2n, log2E = n.bit_length() - 1, E.bit_length() - 1
    ps = codec.ProofStream()
    ps.alter_fiat_shamir_state_with(claim.encode())
    ps.enqueue("Log2PaddedHeight", d["padded_height"].bit_length() - 1)

    def sub(a, step):
        return a if step == 1 else np.ascontiguousarray(a[:, ::step])

    # main table (stark.rs:359-374)
    main_m = corc.to_mont(main_trace).reshape(NM, n)
    mrand_m = corc.to_mont(np.ascontiguousarray(main_rand, dtype=np.uint64)).reshape(NM, h)
    main_lde = corc.lde_table(main_m, mrand_m, off, log2E, mont_io=True)
    clk.mark("LDE (main)")
    main_nodes = corc.merkle_build(corc.hash_rows_colmajor(sub(main_lde, es), mont_io=True), mont_io=True)
    clk.mark("Merkle tree (main)")
    ps.enqueue("MerkleRoot", _digest(main_nodes))
    challenges = S.derive_challenges(ps.sample_scalars(S.NUM_SAMPLED_CHALLENGES), claim)

    # aux table
    aux_trace, aux_rand = aux_provider(challenges)
    clk.mark("aux table (caller)")
    aux_trace = np.ascontiguousarray(aux_trace, dtype=np.uint64).reshape(NA, n, 3)
    aux_rand = np.ascontiguousarray(aux_rand, dtype=np.uint64).reshape(NA, h, 3)
    aux_m = corc.to_mont(np.ascontiguousarray(aux_trace.transpose(0, 2, 1))).reshape(3 * NA, n)
    arand_m = corc.to_mont(np.ascontiguousarray(aux_rand.transpose(0, 2, 1))).reshape(3 * NA, h)
    aux_lde = corc.lde_table(aux_m, arand_m, off, log2E, mont_io=True)
    clk.mark("LDE (aux)")
    aux_nodes = corc.merkle_build(corc.hash_rows_colmajor(sub(aux_lde, es), mont_io=True), mont_io=True)
    clk.mark("Merkle tree (aux)")
    ps.enqueue("MerkleRoot", _digest(aux_nodes))

    # quotient (stark.rs:721-798, 1224-1356)
    w0 = ps.sample_scalars(1)[0]
    num_constraints = sum(len(v) for v in S.constraint_degrees().values())
    quot_weights = S.xpows(w0, num_constraints)
    quot_cw = corc.air_quotient(sub(main_lde, qs), sub(aux_lde[:270], qs), log2n, off, challenges, quot_weights, mont_io=True)
    clk.mark("quotient codeword (AIR)")
    quot_poly = corc.xcoset_interpolate_m(quot_cw, off)
    polys = [np.ascontiguousarray(quot_poly[s::S.NUM_QUOTIENT_SEGMENTS]) for s in range(S.NUM_QUOTIENT_SEGMENTS)]
    qr = np.ascontiguousarray(quot_rand, dtype=np.uint64).reshape(-1, 3)
    assert qr.shape[0] == d["num_quotient_randomizer_coefficients"]
    polys.append(corc.to_mont(qr).reshape(-1, 3))
    zeta_k = pow(S.ZETA, S.NUM_QUOTIENT_SEGMENTS, P)
    for i in range(S.NUM_QUOTIENT_SEGMENTS - 1, -1, -1):
        nxt = polys[i + 1]
        m = max(polys[i].shape[0], nxt.shape[0])
        s_ = np.zeros((m, 3), dtype=np.uint64)
        s_[:polys[i].shape[0]] = polys[i]
        corc.xpoly_add_scaled_arg_m(s_, nxt, (-pow(S.ZETA, i, P)) % P, zeta_k)
        polys[i] = s_
    seg_cw = [corc.xcoset_evaluate_m(p, off, N) for p in polys]
    clk.mark("quotient segments (LDE)")
    seg_table = np.concatenate([np.ascontiguousarray(c.T) for c in seg_cw], axis=0)
    quot_nodes = corc.merkle_build(corc.hash_rows_colmajor(np.ascontiguousarray(seg_table), mont_io=True), mont_io=True)
    clk.mark("Merkle tree (quotient)")
    ps.enqueue("MerkleRoot", _digest(quot_nodes))

    # out-of-domain rows (master_table.rs:348-390, stark.rs:450-495)
    alpha = ps.sample_scalars(1)[0]
    omega = F.primitive_root_of_unity(n)
    alpha_next = F.xscale(alpha, omega)
    rows = {}
    for name, pt in (("cur", alpha), ("next", alpha_next)):
        dods, di = corc.bary_weights_m(log2n, pt)
        rows[name] = (corc.ood_row_m(main_m, 1, dods, di, mrand_m, pt), corc.ood_row_m(aux_m, 3, dods, di, arand_m, pt))
    ps.enqueue("OutOfDomainMainRow", rows["cur"][0])
    ps.enqueue("OutOfDomainAuxRow", rows["cur"][1])
    ps.enqueue("OutOfDomainMainRow", rows["next"][0])
    ps.enqueue("OutOfDomainAuxRow", rows["next"][1])
    alpha_pow = F.xpow(alpha, S.NUM_QUOTIENT_SEGMENTS)
    alpha_zeta_pow = F.xpow(F.xscale(alpha, S.ZETA), S.NUM_QUOTIENT_SEGMENTS)
    ps.enqueue("OutOfDomainQuotientSegments", [corc.xpoly_eval_m(p, alpha_pow) for p in polys[:-1]])
    ps.enqueue("OutOfDomainQuotientSegments", [corc.xpoly_eval_m(p, alpha_zeta_pow) for p in polys[1:]])
    clk.mark("out-of-domain rows")

    # combination codeword + DEEP (stark.rs:498-639)
    wm, wq, wd = ps.sample_scalars(3)
    w_main_aux = S.xpows(wm, NM + NA)
    w_quot = S.xpows(wq, S.NUM_RANDOMIZED_QUOTIENT_SEGMENTS)
    w_deep = S.xpows(wd, S.NUM_DEEP_CODEWORD_COMPONENTS)
    comb = corc.weighted_colsum_m(corc.interpolants_table_m(main_m, mrand_m), corc.interpolants_table_m(aux_m, arand_m), w_main_aux)
    main_aux_cw = corc.xcoset_evaluate_m(comb, off, N)
    mlen = max(p.shape[0] for p in polys)
    shared = np.zeros((mlen, 3), dtype=np.uint64)
    for p, w in zip(polys[1:-1], w_quot[1:-1]):
        corc.xpoly_axpy_m(shared, p, w)
    poly_p, poly_r = shared.copy(), shared.copy()
    corc.xpoly_axpy_m(poly_p, polys[0], w_quot[0])
    corc.xpoly_axpy_m(poly_r, polys[-1], w_quot[-1])
    cw_p, cw_r = corc.xcoset_evaluate_m(poly_p, off, N), corc.xcoset_evaluate_m(poly_r, off, N)
    points = [alpha, alpha_next, alpha_pow, alpha_zeta_pow]
    values = [corc.xpoly_eval_m(comb, alpha), corc.xpoly_eval_m(comb, alpha_next), corc.xpoly_eval_m(poly_p, alpha_pow),
              corc.xpoly_eval_m(poly_r, alpha_zeta_pow)]
    combination = corc.deep_combination_m(main_aux_cw, cw_p, cw_r, off, points, values, w_deep)
    clk.mark("linear combination + DEEP")

    if d["ldt"] == "stir":
        revealed = stir_prove(ps, combination, d["stir"])
    else:
        revealed = fri_prove(ps, combination, d)
    clk.mark("low-degree test")

    if alpha_pow[1] == 0 and alpha_pow[2] == 0:                         # zero-knowledge guard (stark.rs:648-663)
        g = F.primitive_root_of_unity(N)
        pts = {off * pow(g, i, P) % P for i in revealed}
        if alpha_pow[0] in pts or alpha_pow[0] * pow(S.ZETA, S.NUM_QUOTIENT_SEGMENTS, P) % P in pts:
            raise ValueError("ZeroKnowledgeViolation")

    ridx = np.array(revealed)
    main_rows = _can(main_lde[:, ridx * es]).reshape(NM, -1)
    aux_rows = _can(aux_lde[:, ridx * es]).reshape(3 * NA, -1)
    ps.enqueue("MasterMainTableRows", [[int(v) for v in main_rows[:, k]] for k in range(len(revealed))])
    ps.enqueue("AuthenticationStructure", _auth(main_nodes, N, revealed))
    ps.enqueue("MasterAuxTableRows", [[tuple(int(v) for v in aux_rows[3 * q:3 * q + 3, k]) for q in range(NA)] for k in range(len(revealed))])
    ps.enqueue("AuthenticationStructure", _auth(aux_nodes, N, revealed))
    seg_rows = [_can(c[ridx]).reshape(-1, 3) for c in seg_cw]
    ps.enqueue("QuotientSegmentsElements", [[tuple(int(v) for v in seg_rows[s][k]) for s in range(5)] for k in range(len(revealed))])
    ps.enqueue("AuthenticationStructure", _auth(quot_nodes, N, revealed))
    clk.mark("open trace leafs")
    return ps.encode()

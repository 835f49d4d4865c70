"""STIR low-degree test: CPU restatement of triton-vm/src/low_degree_test/stir.rs (+ mod.rs:212-300).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the product path never imports this.

Parity status: the parameter maths is pinned against the figures the reference states in-tree
(`log2_binomial_coefficient` table stir.rs:1542-1580; the worked examples "lambda=160, u=2^23, k=160 ->
n=184" and "u=2^8 -> n=610" in the comment stir.rs:738-746; `stacking` 1528-1539;
`folding_polynomial_gives_expected_coefficients` 1634-1644).  The prover itself has no golden vector in
the reference (its tests are prove-then-verify properties), so it is pinned by the restated verifier
accepting its proofs and rejecting corrupted ones (no reference vector for STIR proofs; the encoding layer and
every primitive it shares with FRI are pinned by the reference's whole-proof digests, tests/test_golden.py).

Polynomials are lists of X-field coefficients (tuples of 3 ints, canonical), little-endian.
"""
import math

import numpy as np

from . import corc, field as F, merkle, tip5
from .field import P

LOG2_FIELD_SIZE_U32 = 8 * 8 * 3          # StirParameters::LOG2_FIELD_SIZE (stir.rs:412-413)
LOG2_DOMAIN_SHRINKAGE = 1                # stir.rs:422


# ---- Reed-Solomon code parameters (mod.rs:212-300) ------------------------------------------------
RS_LOG2_FIELD_SIZE = 191.99999999899228  # ReedSolomonCode::LOG2_FIELD_SIZE (mod.rs:227)


def rs_rate(log2_expansion):
    return 1.0 / float(1 << log2_expansion)


def rs_q_ary_entropy(log2_expansion):                    # mod.rs:258-264
    rate = rs_rate(log2_expansion)
    rate_log_rate = rate * -float(log2_expansion)
    one_m = (1.0 - rate) * math.log2(1.0 - rate)
    return rate - (rate_log_rate + one_m) / RS_LOG2_FIELD_SIZE


def rs_proximity_margin(log2_expansion, soundness="proven"):
    return math.sqrt(rs_rate(log2_expansion)) if soundness == "proven" else rs_q_ary_entropy(log2_expansion)


def rs_slackness(log2_expansion, soundness="proven"):
    return rs_proximity_margin(log2_expansion, soundness) / 20.0


def rs_proximity_parameter(log2_expansion, soundness="proven"):
    return 1.0 - rs_proximity_margin(log2_expansion, soundness) - rs_slackness(log2_expansion, soundness)


def rs_log2_list_size(log2_expansion, soundness="proven", log2_poly_degree=0):     # mod.rs:274-287
    if soundness == "proven":
        ls = 1.0 / (2.0 * math.sqrt(rs_rate(log2_expansion)) * rs_slackness(log2_expansion, soundness))
    else:
        ls = math.pow(2.0, float(log2_poly_degree)) / (rs_q_ary_entropy(log2_expansion) * rs_slackness(log2_expansion, soundness))
    return math.log2(ls)


def log2_binomial_coefficient(a, b):     # stir.rs:854-869 (Kahan-Babuska summation, same order)
    assert a >= b
    log2_binom, compensation = 0.0, 0.0
    for i in range(min(b, a - b)):
        summand = math.log2(float(a - i)) - math.log2(float(i + 1))
        corrected = summand - compensation
        nxt = log2_binom + corrected
        compensation = (nxt - log2_binom) - corrected
        log2_binom = nxt
    return log2_binom


def num_unique_in_domain_queries(security, log2_expansion, soundness="proven"):      # stir.rs:633-639
    return int(math.ceil(-float(security) / math.log2(1.0 - rs_proximity_parameter(log2_expansion, soundness))))


def num_total_in_domain_queries(security, log2_domain_len, num_uniques):   # stir.rs:758-776
    k_minus_1 = num_uniques - 1
    assert k_minus_1 >= 0
    domain_len = 1 << log2_domain_len
    l = min(k_minus_1, domain_len // 2)
    log2_u_choose_l = log2_binomial_coefficient(domain_len, l)
    log2_k_minus_1 = max(math.log2(float(k_minus_1)), 0.0) if k_minus_1 > 0 else 0.0
    n = (float(security) + log2_k_minus_1 + log2_u_choose_l) / (float(log2_domain_len) - log2_k_minus_1)
    return int(math.ceil(n))


def num_in_domain_queries(security, log2_domain_size, log2_expansion, soundness="proven"):     # stir.rs:597-609
    uniques = min(num_unique_in_domain_queries(security, log2_expansion, soundness), 1 << log2_domain_size)
    return num_total_in_domain_queries(security, log2_domain_size, uniques)


def num_ood_queries(security, log2_poly_degree, log2_expansion, soundness="proven"):            # stir.rs:831-842
    return int(math.ceil((float(security) - 1.0 + 2.0 * rs_log2_list_size(log2_expansion, soundness, log2_poly_degree))
                         / float(LOG2_FIELD_SIZE_U32 - log2_poly_degree)))


def derive(security, log2_folding_factor, log2_initial_expansion, log2_high_degree_bound, soundness="proven"):
    """StirParameters::try_into_stir (stir.rs:437-567) -> dict"""
    if log2_folding_factor < 2: raise ValueError("TooSmallLog2FoldingFactor")
    if log2_initial_expansion == 0: raise ValueError("TooSmallInitialExpansionFactor")
    if log2_high_degree_bound < log2_folding_factor: raise ValueError("TooLowDegreeOfHighDegreePolynomials")
    folding_factor = 1 << log2_folding_factor
    folded_poly_degree = ((1 << log2_high_degree_bound) - 1) // folding_factor
    log2_expansion = log2_initial_expansion
    log2_domain_len = log2_high_degree_bound + log2_initial_expansion
    if log2_domain_len > 32: raise ValueError("InitialDomainTooBig")
    log2_folded_domain_size = log2_domain_len - log2_folding_factor
    rounds = []
    while folded_poly_degree > folding_factor:
        in_domain = num_in_domain_queries(security, log2_folded_domain_size, log2_expansion, soundness)
        log2_next_expansion = log2_expansion + log2_folding_factor - LOG2_DOMAIN_SHRINKAGE
        ood = num_ood_queries(security, folded_poly_degree.bit_length() - 1, log2_next_expansion, soundness)
        next_deg = folded_poly_degree // folding_factor
        if in_domain + ood > next_deg:
            break
        rounds.append((in_domain, ood))
        folded_poly_degree = next_deg
        log2_expansion = log2_next_expansion
        log2_folded_domain_size -= LOG2_DOMAIN_SHRINKAGE
    final_in = num_in_domain_queries(security, log2_folded_domain_size, log2_expansion, soundness)
    return dict(initial_domain_len=1 << log2_domain_len, initial_offset=F.GENERATOR, folding_factor=folding_factor,
                round_queries=rounds, final_num_in_domain_queries=final_in, final_degree=folded_poly_degree,
                num_first_round_queries=rounds[0][0] if rounds else final_in)


# ---- polynomial helpers ----------------------------------------------------------------------------
def ptrim(c):
    c = list(c)
    while c and c[-1] == F.X_ZERO:
        c.pop()
    return c


def peval(c, x):
    acc = F.X_ZERO
    for v in reversed(c):
        acc = F.xadd(F.xmul(acc, x), v)
    return acc


def pmul(a, b):
    if not a or not b: return []
    out = [F.X_ZERO] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        if x == F.X_ZERO: continue
        for j, y in enumerate(b):
            out[i + j] = F.xadd(out[i + j], F.xmul(x, y))
    return out


def psub(a, b):
    n = max(len(a), len(b))
    a = list(a) + [F.X_ZERO] * (n - len(a)); b = list(b) + [F.X_ZERO] * (n - len(b))
    return ptrim([F.xsub(x, y) for x, y in zip(a, b)])


def zerofier(points):                    # Polynomial::zerofier
    z = [F.X_ONE]
    for p_ in points:
        z = pmul(z, [F.xneg(p_), F.X_ONE])
    return z


def interpolate(xs, ys):                 # Polynomial::interpolate (the unique interpolant; Lagrange form)
    z = zerofier(xs)
    out = [F.X_ZERO] * len(xs)
    for xi, yi in zip(xs, ys):
        # z / (X - xi) by synthetic division, then scale by yi / prod_{j != i}(xi - xj)
        q, acc = [F.X_ZERO] * (len(z) - 1), F.X_ZERO
        for k in range(len(z) - 1, 0, -1):
            acc = F.xadd(z[k], F.xmul(acc, xi))
            q[k - 1] = acc
        denom = peval(q, xi)
        s = F.xmul(yi, F.xinv(denom))
        out = [F.xadd(o, F.xmul(s, c)) for o, c in zip(out, q)]
    return ptrim(out)


def pdiv_exact(num, den):                # (a - b) / zerofier; the reference's `/` is Euclidean division, remainder dropped
    num, den = ptrim(num), ptrim(den)
    if len(num) < len(den): return []
    lead_inv = F.xinv(den[-1])
    num = list(num)
    q = [F.X_ZERO] * (len(num) - len(den) + 1)
    for k in range(len(q) - 1, -1, -1):
        c = F.xmul(num[k + len(den) - 1], lead_inv)
        q[k] = c
        if c != F.X_ZERO:
            for j, dv in enumerate(den):
                num[k + j] = F.xsub(num[k + j], F.xmul(c, dv))
    return ptrim(q)


def fold_polynomial(c, folding_factor, r):          # stir.rs:1132-1147
    return ptrim([peval(c[i:i + folding_factor], r) for i in range(0, len(c), folding_factor)])


def xevaluate(coeffs, offset, n):
    """ArithmeticDomain::evaluate for an X-field polynomial of any length (chunked like arithmetic_domain.rs:153-167)"""
    c = np.array([list(v) for v in coeffs], dtype=np.uint64).reshape(-1, 3)
    out = np.zeros((n, 3), dtype=np.uint64)
    if c.shape[0] == 0:
        return out
    # reduce modulo X^n - offset^n (the domain's zerofier) - evaluations are unchanged
    if c.shape[0] > n:
        on = pow(offset, n, P)
        red = [F.X_ZERO] * n
        scale = 1
        for s in range(0, c.shape[0], n):
            for j, v in enumerate(c[s:s + n]):
                red[j] = F.xadd(red[j], F.xscale(tuple(int(t) for t in v), scale))
            scale = scale * on % P
        c = np.array([list(v) for v in red], dtype=np.uint64)
    log2n = n.bit_length() - 1
    return np.stack([corc.coset_evaluate(np.ascontiguousarray(c[:, d]), offset, log2n) for d in range(3)], axis=1)


def xinterpolate(values, offset):
    v = np.array(values, dtype=np.uint64).reshape(-1, 3)
    co = np.stack([corc.coset_interpolate(np.ascontiguousarray(v[:, d]), offset) for d in range(3)], axis=1)
    return ptrim([tuple(int(t) for t in row) for row in co])


def domain_value(offset, length, i):
    return offset * pow(F.primitive_root_of_unity(length), i, P) % P


def next_round_domain(offset, length):             # stir.rs:1149-1155: pow(2), then offset *= old offset
    return pow(offset, 3, P), length // 2


def unique(seq):
    seen, out = set(), []
    for v in seq:
        if v not in seen:
            seen.add(v); out.append(v)
    return out


# ---- stacked Merkle tree (stir.rs:1374-1433) --------------------------------------------------------
def stack(codeword, stack_height):
    dist = -(-len(codeword) // stack_height)
    return [[codeword[j] for j in range(skip, len(codeword), dist)] for skip in range(dist)]


class StirMerkleTree:
    def __init__(self, codeword, stack_height):
        cw = [tuple(int(t) for t in v) for v in codeword]
        self.stacked = stack(cw, stack_height)
        flat = np.array([[t for x in st for t in x] for st in self.stacked], dtype=np.uint64)   # bfe_slice of each stack
        digests = corc.hash_rows_colmajor(np.ascontiguousarray(flat.T))
        self.nodes = corc.merkle_build(digests)

    def root(self):
        return [int(v) for v in self.nodes[1]]

    def inclusion_proof(self, indices):
        leafs = [self.stacked[i] for i in indices]
        n = len(self.stacked)
        auth = [[int(v) for v in self.nodes[k]] for k in merkle.auth_structure_node_indices(n, indices)]
        return leafs, auth


# ---- prover (stir.rs:885-993) ---------------------------------------------------------------------------
def prove(ps, codeword, sp):
    """codeword: [N][3] evaluations on the initial domain.  Enqueues into `ps`; returns the revealed
    first-round indices."""
    offset, length = sp["initial_offset"], sp["initial_domain_len"]
    assert len(codeword) == length
    ff = sp["folding_factor"]
    commitment = StirMerkleTree(codeword, ff)
    ps.enqueue("MerkleRoot", commitment.root())
    poly = xinterpolate(codeword, offset)
    first_round_indices = None
    for in_domain, ood in sp["round_queries"]:
        r = ps.sample_scalars(1)[0]
        folded = fold_polynomial(poly, ff, r)
        n_off, n_len = next_round_domain(offset, length)
        folded_evals = xevaluate(folded, n_off, n_len)
        folded_commitment = StirMerkleTree(folded_evals, ff)
        ps.enqueue("MerkleRoot", folded_commitment.root())
        ood_queries = ps.sample_scalars(ood)
        ood_values = [peval(folded, x) for x in ood_queries]
        ps.enqueue("StirOutOfDomainValues", ood_values)
        queried = ps.sample_indices(length, in_domain)
        f_off, f_len = pow(offset, ff, P), length // ff          # domain.pow(folding_factor)
        folded_idx = unique([i % f_len for i in queried])
        ps.enqueue("StirResponse", commitment.inclusion_proof(folded_idx))
        qvals = [domain_value(f_off, f_len, i) for i in folded_idx]
        points = [F.xlift(x) for x in qvals] + list(ood_queries)
        answers = [peval(folded, F.xlift(x)) for x in qvals] + ood_values
        ans = interpolate(points, answers)
        quotient = pdiv_exact(psub(folded, ans), zerofier(points))
        dcr = ps.sample_scalars(1)[0]
        dc = []
        acc = F.X_ONE
        for _ in range(len(answers) + 1):
            dc.append(acc); acc = F.xmul(acc, dcr)
        poly = ptrim(pmul(quotient, dc))
        offset, length, commitment = n_off, n_len, folded_commitment
        if first_round_indices is None:
            first_round_indices = queried
    r = ps.sample_scalars(1)[0]
    final_poly = fold_polynomial(poly, ff, r)
    ps.enqueue("Polynomial", final_poly)
    f_len = length // ff
    queried = ps.sample_indices(length, sp["final_num_in_domain_queries"])
    ps.enqueue("StirResponse", commitment.inclusion_proof(unique([i % f_len for i in queried])))
    return first_round_indices if first_round_indices is not None else queried


# ---- verifier (stir.rs:995-1108, 1157-1323, 1435-1468) -----------------------------------------------------
def _extract(ps, sp, offset, length, num_queries):
    ff = sp["folding_factor"]
    queried = ps.sample_indices(length, num_queries)
    leafs, auth = ps.dequeue("StirResponse")
    f_off, f_len = pow(offset, ff, P), length // ff
    folded_idx = unique([i % f_len for i in queried])
    if len(leafs) != len(folded_idx):
        raise ValueError("IncorrectNumberOfRevealedLeaves")
    by_idx = dict(zip(folded_idx, leafs))
    g = F.primitive_root_of_unity(length)
    kth = pow(g, f_len, P)
    queries = []
    for index in queried:
        qi = index % f_len
        queries.append(dict(index=index, fidx=qi, point=domain_value(f_off, f_len, qi), root=offset * pow(g, qi, P) % P,
                            kth=kth, values=by_idx[qi]))
    return queries, auth, f_len


def _authenticate(queries, auth, f_len, root):
    indexed = {}
    for q in queries:
        words = [t for x in q["values"] for t in x]
        indexed[q["fidx"]] = [int(v) for v in tip5.hash_varlen(words)]
    if not merkle.verify_inclusion(root, f_len.bit_length() - 1, sorted(indexed.items()), auth):
        raise ValueError("BadMerkleAuthenticationPath")


def _coset_interp_eval(root, values, x):
    """Polynomial::fast_coset_interpolate(root, values).evaluate(x)"""
    k = len(values)
    co = [[int(t) for t in corc.coset_interpolate(np.array([v[d] for v in values], dtype=np.uint64), root)] for d in range(3)]
    return peval([(co[0][j], co[1][j], co[2][j]) for j in range(k)], x)


def _initial_answers(queries, r):
    return [_coset_interp_eval(q["root"], q["values"], r) for q in queries]


def _subsequent_answers(qd, queries, r):
    quotient_set, quotient_answers, dcr = qd
    ans = interpolate(quotient_set, quotient_answers)
    zf = zerofier(quotient_set)
    degree_difference = len(quotient_set) + 1
    out = []
    for q in queries:
        cur, evals = q["root"], []
        for ev in q["values"]:
            xc = F.xlift(cur)
            quot = F.xmul(F.xsub(ev, peval(ans, xc)), F.xinv(peval(zf, xc)))
            common = F.xscale(dcr, cur)
            if common == F.X_ONE:
                dcf = F.xlift(degree_difference)
            else:
                dcf = F.xmul(F.xsub(F.X_ONE, F.xpow(common, degree_difference)), F.xinv(F.xsub(F.X_ONE, common)))
            evals.append(F.xmul(dcf, quot))
            cur = cur * q["kth"] % P
        out.append(_coset_interp_eval(q["root"], evals, r))
    return out


def verify(ps, sp):
    """-> (first_round_indices, partial_first_codeword)"""
    ff = sp["folding_factor"]
    offset, length = sp["initial_offset"], sp["initial_domain_len"]
    prev_root = ps.dequeue("MerkleRoot")
    partial_first, first_indices, qd = None, None, None
    for in_domain, ood in sp["round_queries"]:
        r = ps.sample_scalars(1)[0]
        cur_root = ps.dequeue("MerkleRoot")
        ood_queries = ps.sample_scalars(ood)
        ood_answers = ps.dequeue("StirOutOfDomainValues")
        queries, auth, f_len = _extract(ps, sp, offset, length, in_domain)
        _authenticate(queries, auth, f_len, prev_root)
        if partial_first is None:
            partial_first = [q["values"][q["index"] // f_len] for q in queries]
            first_indices = [q["index"] for q in queries]
        answers = _initial_answers(queries, r) if qd is None else _subsequent_answers(qd, queries, r)
        seen, qs, qa = set(), [], []
        for pt, a in list(zip([F.xlift(q["point"]) for q in queries], answers)) + list(zip(ood_queries, ood_answers)):
            if pt not in seen:
                seen.add(pt); qs.append(pt); qa.append(a)
        dcr = ps.sample_scalars(1)[0]
        qd = (qs, qa, dcr)
        offset, length = next_round_domain(offset, length)
        prev_root = cur_root
    r = ps.sample_scalars(1)[0]
    poly = ps.dequeue("Polynomial")
    if max(len(poly) - 1, 0) > sp["final_degree"]:
        raise ValueError("LastRoundPolynomialHasTooHighDegree")
    queries, auth, f_len = _extract(ps, sp, offset, length, sp["final_num_in_domain_queries"])
    _authenticate(queries, auth, f_len, prev_root)
    final_answers = _initial_answers(queries, r) if qd is None else _subsequent_answers(qd, queries, r)
    for q, a in zip(queries, final_answers):
        if peval(poly, F.xlift(q["point"])) != a:
            raise ValueError("LastRoundPolynomialEvaluationMismatch")
    if partial_first is None:
        partial_first = [q["values"][q["index"] // f_len] for q in queries]
        first_indices = [q["index"] for q in queries]
    return first_indices, partial_first

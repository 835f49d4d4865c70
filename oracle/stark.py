"""CPU restatement of `Stark::prove` / `Stark::verify` with the FRI or STIR low-degree test.

Follows the reference file by file:
  parameters      triton-vm/src/stark.rs:1885-2089, low_degree_test/mod.rs:215-360, fri.rs:799-924
  domains         stark.rs:263-286, arithmetic_domain.rs
  prover          stark.rs:331-719 (+721-798 cached quotient path, 1224-1379 helpers)
  master tables   table/master_table.rs:258-609, 1194-1363
  FRI             low_degree_test/fri.rs:212-366, 393-735, 754-772
  STIR            low_degree_test/stir.rs (oracle/stir.py)
  verifier        stark.rs:1388-1763
  transcript      oracle/codec.py

Randomness stays on the caller's side of the boundary (SURVEY.md §7 hard part 1): the trace
randomizer coefficients and the quotient-segment randomizer are inputs.

Heavy loops run in the C oracle (oracle/c) — still a plain sequential restatement.  Tables are
numpy uint64 arrays in canonical form: main [379, n]; aux [91, n, 3]; X-field codewords [N, 3].

`Stark(ldt=None)` follows the reference's heuristic (`Stark::default()`: FRI below padded height 2^16,
STIR from there on, stark.rs:1942-1958); `ldt="fri"` / `"stir"` is `with_ldt_choice`.

TEST INFRASTRUCTURE ONLY."""
import math

import numpy as np

from . import codec, corc, field as F, merkle, stir as stir_mod, tip5
from .field import P

NUM_MAIN_COLUMNS = 379
NUM_AUX_COLUMNS = 91
NUM_QUOTIENT_SEGMENTS = 4
NUM_RANDOMIZED_QUOTIENT_SEGMENTS = 5
NUM_DEEP_CODEWORD_COMPONENTS = 4
AIR_FAN_IN = 2
NUM_OUT_OF_DOMAIN_QUOTIENTS = 1
ZETA = 3                         # stark.rs:1801
NUM_SAMPLED_CHALLENGES = 59
CH_COMPRESS_PROGRAM_DIGEST, CH_STD_IN, CH_STD_OUT = 0, 1, 2     # challenge_id.rs order
CH_LOOKUP_PUBLIC_INDETERMINATE = 54
CURRENT_VERSION = 6              # proof.rs:33


def next_pow2(x):
    return 1 if x <= 1 else 1 << (x - 1).bit_length()


# ---- AIR metadata (constraint degrees, number of constraints) -----------------------------------
_AIR = None


def air():
    global _AIR
    if _AIR is None:
        import os, sys
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "triton-vm_b200"))
        from airgen.build import build_air
        _AIR = build_air()
    return _AIR


def constraint_degrees():
    a = air()
    return {cat: [a.builders[cat].degree(c) for c in a.constraints[cat]] for cat in ("init", "cons", "tran", "term")}


# ---- parameters ------------------------------------------------------------------------------------
class Stark:
    """stark.rs:113-145 — (security_level, log2 expansion factor, LDT choice); proven regime.
    `ldt`: "fri", "stir" or None = the reference's heuristic (stark.rs:1942-1957: FRI below padded
    height 2^16, STIR from there on)."""

    def __init__(self, security_level=160, log2_expansion_factor=2, ldt=None, soundness="proven"):
        assert log2_expansion_factor >= 1
        assert ldt in (None, "fri", "stir") and soundness in ("proven", "conjectured")
        self.security_level = security_level
        self.log2_expansion = log2_expansion_factor
        self.ldt = ldt
        self.soundness = soundness                                # ProximityRegime (mod.rs:60-80)

    def ldt_choice(self, padded_height):
        threshold = 16 if self.soundness == "proven" else 17      # stark.rs:1944-1949
        return self.ldt or ("fri" if next_pow2(padded_height).bit_length() - 1 < threshold else "stir")

    # low_degree_test/mod.rs:250-300 (ProximityRegime::Proven), fri.rs:832-836
    def num_collinearity_checks(self):
        proximity_parameter = stir_mod.rs_proximity_parameter(self.log2_expansion, self.soundness)
        return int(math.ceil(-float(self.security_level) / math.log2(1.0 - proximity_parameter)))

    @staticmethod
    def num_trace_randomizers_for(num_first_round_queries):      # stark.rs:2083-2089
        return num_first_round_queries + NUM_QUOTIENT_SEGMENTS * 3 * AIR_FAN_IN + 1

    @staticmethod
    def num_quotient_table_randomizers(h):                        # stark.rs:1894-1896
        return (h + NUM_OUT_OF_DOMAIN_QUOTIENTS) * NUM_RANDOMIZED_QUOTIENT_SEGMENTS

    @staticmethod
    def randomized_trace_len(padded_height, h):                   # stark.rs:1885-1890
        return next_pow2(max(padded_height + h, 2 * h + 1, Stark.num_quotient_table_randomizers(h)))

    def derive(self, padded_height):
        """-> dict of every derived quantity (stark.rs:342-353, 1905-2075, 263-286; fri.rs:816-924)"""
        padded_height = next_pow2(padded_height)
        log2_ph = padded_height.bit_length() - 1
        checks = self.num_collinearity_checks()
        expansion = 1 << self.log2_expansion
        ldt = self.ldt_choice(padded_height)
        stir_params = None
        hdb = log2_ph
        while True:                                               # stark.rs:1972-2060
            hdb += 1
            ldt_len = 1 << (hdb + self.log2_expansion)
            if ldt == "stir":
                stir_params = stir_mod.derive(self.security_level, 2, self.log2_expansion, hdb, self.soundness)
                first_round = stir_params["num_first_round_queries"]
            else:
                first_round = checks
            h = self.num_trace_randomizers_for(first_round)
            rtl = self.randomized_trace_len(padded_height, h)
            if ldt_len >= rtl * expansion:
                break
        interpolant_degree = rtl - 1
        deg = constraint_degrees()
        zerofier_degree = {"init": 1, "cons": padded_height, "tran": padded_height - 1, "term": 1}
        max_constraint_degree = max(interpolant_degree * d - zerofier_degree[cat]
                                    for cat, ds in deg.items() for d in ds)   # auxiliary_table.rs / codegen.rs:222-229
        max_degree = next_pow2(max_constraint_degree) - 1         # stark.rs:1905-1916
        quotient_len = next_pow2(max_degree)
        fri_max_degree = ldt_len // expansion - 1                 # fri.rs:885-887
        max_num_rounds = next_pow2(fri_max_degree + 1).bit_length() - 1
        skip = (checks.bit_length() - 1 if checks else 0) + 1     # fri.rs:907-920
        num_rounds = max(0, max_num_rounds - skip)
        return dict(padded_height=padded_height, num_trace_randomizers=h, randomized_trace_len=rtl,
                    trace_len=rtl // 2, quotient_len=quotient_len, ldt_len=ldt_len, ldt_offset=F.GENERATOR,
                    num_collinearity_checks=checks, fri_num_rounds=num_rounds,
                    fri_last_round_max_degree=fri_max_degree >> num_rounds,
                    num_quotient_randomizer_coefficients=self.num_quotient_table_randomizers(h), max_degree=max_degree,
                    ldt=ldt, stir=stir_params, num_first_round_queries=first_round)


class Claim:
    def __init__(self, program_digest, inp=(), out=(), version=CURRENT_VERSION):
        self.program_digest, self.input, self.output, self.version = list(program_digest), list(inp), list(out), version

    def encode(self):
        return codec.encode_claim(self.program_digest, self.version, self.input, self.output)


# ---- small helpers -------------------------------------------------------------------------------
def xpows(x, n, start=0):
    out, acc = [], F.xpow(x, start)
    for _ in range(n):
        out.append(acc)
        acc = F.xmul(acc, x)
    return out


def eval_arg_terminal(symbols, initial, challenge):              # cross_table_argument.rs:60-73
    acc = initial
    for s in symbols:
        acc = F.xadd(F.xmul(challenge, acc), (int(s) % P, 0, 0))
    return acc


def derive_challenges(sampled, claim):                           # challenges.rs:88-135
    ch = list(sampled)
    assert len(ch) == NUM_SAMPLED_CHALLENGES
    compressed_digest = eval_arg_terminal(claim.program_digest, F.X_ONE, ch[CH_COMPRESS_PROGRAM_DIGEST])
    input_terminal = eval_arg_terminal(claim.input, F.X_ONE, ch[CH_STD_IN])
    output_terminal = eval_arg_terminal(claim.output, F.X_ONE, ch[CH_STD_OUT])
    lookup_terminal = eval_arg_terminal(tip5.LOOKUP_TABLE, F.X_ONE, ch[CH_LOOKUP_PUBLIC_INDETERMINATE])
    return ch + [input_terminal, output_terminal, lookup_terminal, compressed_digest]


def domain_values(offset, n):
    g = F.primitive_root_of_unity(n)
    out, acc = [], offset % P
    for _ in range(n):
        out.append(acc)
        acc = acc * g % P
    return out


def xpoly_eval(coeffs, x):
    acc = F.X_ZERO
    for c in reversed(coeffs):
        acc = F.xadd(F.xmul(acc, x), tuple(int(v) for v in c))
    return acc


def as_x(a):
    return [tuple(int(v) for v in row) for row in a]


def xcoset_evaluate(coeffs, offset, n):
    """ArithmeticDomain::evaluate on an X-field polynomial; coeffs [k,3] -> [n,3]"""
    c = np.array(coeffs, dtype=np.uint64).reshape(-1, 3)
    if c.shape[0] == 0:
        return np.zeros((n, 3), dtype=np.uint64)
    return np.stack([corc.coset_evaluate(np.ascontiguousarray(c[:, d]), offset, n.bit_length() - 1) for d in range(3)], axis=1)


def xcoset_interpolate(values, offset):
    v = np.array(values, dtype=np.uint64).reshape(-1, 3)
    return np.stack([corc.coset_interpolate(np.ascontiguousarray(v[:, d]), offset) for d in range(3)], axis=1)


def xrow_digests(cols):
    """cols: list of [N,3] arrays (X-field columns) -> Tip5 digest of every row (row = all coefficients)"""
    flat = np.concatenate([np.ascontiguousarray(c.T) for c in cols], axis=0)    # [3*ncols, N], order col0.c0,c1,c2,col1...
    return corc.hash_rows_colmajor(np.ascontiguousarray(flat))


def xfe_leaf_digests(cw):
    """Digest::from(xfe) = (c0, c1, c2, 0, 0): fri.rs:343-347, 927-929"""
    cw = np.array(cw, dtype=np.uint64).reshape(-1, 3)
    return np.concatenate([cw, np.zeros((cw.shape[0], 2), dtype=np.uint64)], axis=1)


# ---- master-table operations ------------------------------------------------------------------------
def column_interpolants(trace, rand):
    """randomized_column_interpolant for every column (master_table.rs:392-403): [ncols, 2n]"""
    ncols, n = trace.shape
    out = np.zeros((ncols, 2 * n), dtype=np.uint64)
    for c in range(ncols):
        co = [int(v) for v in corc.coset_interpolate(trace[c], 1)]
        if rand is not None:
            for i in range(rand.shape[1]):
                r = int(rand[c, i])
                co[i] = (co[i] - r) % P
            co = co + [0] * n
            for i in range(rand.shape[1]):
                co[n + i] = (co[n + i] + int(rand[c, i])) % P
        else:
            co = co + [0] * n
        out[c] = np.array(co, dtype=np.uint64)
    return out


def out_of_domain_row(trace_x, rand_x, alpha):
    """master_table.rs:348-390: batched barycentric evaluation + zerofier(alpha)*randomizer(alpha).
    trace_x: list over columns of lists of X-field values (main columns lifted); rand_x likewise."""
    n = len(trace_x[0])
    domain = domain_values(1, n)
    shift_inv = F.xbatch_inversion([F.xsub(alpha, (d, 0, 0)) for d in domain])
    dods = [F.xscale(inv, d) for d, inv in zip(domain, shift_inv)]
    denom = F.X_ZERO
    for v in dods:
        denom = F.xadd(denom, v)
    denom_inv = F.xinv(denom)
    zerofier = F.xsub(F.xpow(alpha, n), F.X_ONE)                 # trace-domain offset is 1
    out = []
    for col, rnd in zip(trace_x, rand_x):
        num = F.X_ZERO
        for v, d in zip(col, dods):
            num = F.xadd(num, F.xmul(v, d))
        r_at_alpha = xpoly_eval(rnd, alpha)
        out.append(F.xadd(F.xmul(num, denom_inv), F.xmul(zerofier, r_at_alpha)))
    return out


# ---- FRI --------------------------------------------------------------------------------------------
def fri_prove(ps, codeword, d):
    """fri.rs:212-366, 754-772 -> first-round A indices; `codeword` [N,3]"""
    rounds = []
    cw = np.array(codeword, dtype=np.uint64)
    offset, length = d["ldt_offset"], d["ldt_len"]
    nodes = corc.merkle_build(xfe_leaf_digests(cw))
    rounds.append((offset, cw, nodes))
    ps.enqueue("MerkleRoot", [int(v) for v in nodes[1]])
    for _ in range(d["fri_num_rounds"]):
        chal = ps.sample_scalars(1)[0]
        cw = corc.fri_fold(cw, offset, chal)
        offset = offset * offset % P
        nodes = corc.merkle_build(xfe_leaf_digests(cw))
        rounds.append((offset, cw, nodes))
        ps.enqueue("MerkleRoot", [int(v) for v in nodes[1]])
    last = rounds[-1][1]
    ps.enqueue("FriCodeword", as_x(last))
    ps.enqueue("Polynomial", as_x(xcoset_interpolate(last, 1)))
    a_indices = ps.sample_indices(length, d["num_collinearity_checks"])

    def reveal(rnd, idx):
        _, c, nd = rounds[rnd]
        leaves = [tuple(int(v) for v in c[i]) for i in idx]
        n = c.shape[0]
        auth = [[int(v) for v in nd[k]] for k in merkle.auth_structure_node_indices(n, idx)]
        ps.enqueue("FriResponse", (leaves, auth))

    reveal(0, a_indices)
    for rnd in range(len(rounds) - 1):
        n = rounds[rnd][1].shape[0]
        reveal(rnd, [(a + n // 2) % n for a in a_indices])
    ps.sample_scalars(1)                                          # fri.rs:764-769
    return a_indices, rounds


def barycentric_evaluate(codeword, x):
    """twenty-first barycentric_evaluate over the unit-offset domain of len(codeword) (fri.rs:675)"""
    n = len(codeword)
    domain = domain_values(1, n)
    shift_inv = F.xbatch_inversion([F.xsub(x, (d, 0, 0)) for d in domain])
    num, den = F.X_ZERO, F.X_ZERO
    for d, inv, v in zip(domain, shift_inv, codeword):
        w = F.xscale(inv, d)
        num = F.xadd(num, F.xmul(w, v))
        den = F.xadd(den, w)
    return F.xmul(num, F.xinv(den))


def fri_verify(ps, d):
    """fri.rs:393-735 -> (first_round_indices, partial_first_codeword)"""
    checks, num_rounds = d["num_collinearity_checks"], d["fri_num_rounds"]
    rounds = []
    offset, length = d["ldt_offset"], d["ldt_len"]
    for j in range(num_rounds + 1):
        root = ps.dequeue("MerkleRoot")
        chal = ps.sample_scalars(1)[0] if (num_rounds > 0 and j <= num_rounds - 1) else None
        rounds.append(dict(offset=offset, len=length, root=root, chal=chal))
        offset, length = offset * offset % P, length // 2
    last_codeword = ps.dequeue("FriCodeword")
    last_poly = ps.dequeue("Polynomial")
    if len(last_codeword) != rounds[-1]["len"]:
        raise ValueError("LastCodewordMismatch")
    a_indices = ps.sample_indices(d["ldt_len"], checks)

    def check(rnd, idx, leaves, auth):
        if len(leaves) != checks:
            raise ValueError("IncorrectNumberOfRevealedLeaves")
        height = rounds[rnd]["len"].bit_length() - 1
        indexed = [(i, [l[0], l[1], l[2], 0, 0]) for i, l in zip(idx, leaves)]
        if not merkle.verify_inclusion(rounds[rnd]["root"], height, indexed, auth):
            raise ValueError("BadMerkleAuthenticationPath")

    leaves_a, auth_a = ps.dequeue("FriResponse")
    check(0, [a % rounds[0]["len"] for a in a_indices], leaves_a, auth_a)
    rounds[0]["a"] = leaves_a
    for rnd in range(num_rounds):
        n = rounds[rnd]["len"]
        leaves_b, auth_b = ps.dequeue("FriResponse")
        check(rnd, [(a + n // 2) % n for a in a_indices], leaves_b, auth_b)
        rounds[rnd]["b"] = leaves_b
    for rnd in range(num_rounds):                                 # fold (fri.rs:563-584, get_colinear_y)
        r = rounds[rnd]
        n, g = r["len"], F.primitive_root_of_unity(r["len"])
        folded = []
        for i, a in enumerate(a_indices):
            ia, ib = a % n, (a + n // 2) % n
            xa = r["offset"] * pow(g, ia, P) % P
            xb = r["offset"] * pow(g, ib, P) % P
            ya, yb = r["a"][i], r["b"][i]
            # line through (xa, ya), (xb, yb) evaluated at the folding challenge
            slope = F.xscale(F.xsub(yb, ya), F.inv((xb - xa) % P))
            folded.append(F.xadd(ya, F.xmul(slope, F.xsub(r["chal"], (xa, 0, 0)))))
        rounds[rnd + 1]["a"] = folded
    nodes = corc.merkle_build(xfe_leaf_digests(last_codeword))
    if [int(v) for v in nodes[1]] != list(rounds[-1]["root"]):
        raise ValueError("BadMerkleRootForLastCodeword")
    nl = rounds[-1]["len"]
    if [last_codeword[a % nl] for a in a_indices] != rounds[-1]["a"]:
        raise ValueError("LastCodewordMismatch")
    if len(last_poly) - 1 > d["fri_last_round_max_degree"]:
        raise ValueError("LastRoundPolynomialHasTooHighDegree")
    x = ps.sample_scalars(1)[0]
    if xpoly_eval(last_poly, x) != barycentric_evaluate(last_codeword, x):
        raise ValueError("LastRoundPolynomialEvaluationMismatch")
    return a_indices, rounds[0]["a"]


# ---- prover -----------------------------------------------------------------------------------------
def prove(stark, claim, main_trace, main_rand, aux_provider, quot_rand, padded_height=None, keep=False):
    """Prover::prove (stark.rs:331-719).
    main_trace [379, n] (n = derived trace-domain length), main_rand [379, h];
    aux_provider(challenges: list of 63 X-field) -> (aux_trace [91, n, 3], aux_rand [91, h, 3]);
    quot_rand [(h+1)*5, 3].  Returns (proof words, artefacts dict)."""
    main_trace = np.ascontiguousarray(main_trace, dtype=np.uint64)
    n = main_trace.shape[1]
    d = stark.derive(padded_height or n)
    assert d["trace_len"] == n, (d["trace_len"], n)
    # the tables are extended over the larger of the quotient and the LDT domain (evaluation domain, master_table.rs:
    # 258-322); the AIR sees every qs-th row of it (quotient_domain_table, 769-779), commitments and openings every
    # es-th row (ldt_domain_table)
    h, N, off = d["num_trace_randomizers"], d["ldt_len"], d["ldt_offset"]
    E = max(N, d["quotient_len"])
    es, qs = E // N, E // d["quotient_len"]
    log2N = N.bit_length() - 1
    log2E = E.bit_length() - 1
    art = {"derived": d}

    ps = codec.ProofStream()
    ps.alter_fiat_shamir_state_with(claim.encode())
    ps.enqueue("Log2PaddedHeight", d["padded_height"].bit_length() - 1)

    # main table: LDE, row hashes, Merkle tree
    main_rand = np.ascontiguousarray(main_rand, dtype=np.uint64).reshape(NUM_MAIN_COLUMNS, h)
    main_lde = corc.lde_table(main_trace, main_rand, off, log2E)                     # [379, E]
    main_nodes = corc.merkle_build(corc.hash_rows_colmajor(np.ascontiguousarray(main_lde[:, ::es])))
    ps.enqueue("MerkleRoot", [int(v) for v in main_nodes[1]])
    challenges = derive_challenges(ps.sample_scalars(NUM_SAMPLED_CHALLENGES), claim)

    # aux table
    aux_trace, aux_rand = aux_provider(challenges)
    aux_trace = np.ascontiguousarray(aux_trace, dtype=np.uint64).reshape(NUM_AUX_COLUMNS, n, 3)
    aux_rand = np.ascontiguousarray(aux_rand, dtype=np.uint64).reshape(NUM_AUX_COLUMNS, h, 3)
    aux_planar = np.ascontiguousarray(aux_trace.transpose(0, 2, 1)).reshape(3 * NUM_AUX_COLUMNS, n)   # column 3q+d
    aux_rand_planar = np.ascontiguousarray(aux_rand.transpose(0, 2, 1)).reshape(3 * NUM_AUX_COLUMNS, h)
    aux_lde = corc.lde_table(aux_planar, aux_rand_planar, off, log2E)                # [273, E]
    aux_nodes = corc.merkle_build(corc.hash_rows_colmajor(np.ascontiguousarray(aux_lde[:, ::es])))
    ps.enqueue("MerkleRoot", [int(v) for v in aux_nodes[1]])

    # quotient
    w0 = ps.sample_scalars(1)[0]
    num_constraints = sum(len(v) for v in constraint_degrees().values())
    quot_weights = xpows(w0, num_constraints)
    quotient_codeword = corc.air_quotient(np.ascontiguousarray(main_lde[:, ::qs]), np.ascontiguousarray(aux_lde[:270, ::qs]),
                                          n.bit_length() - 1, off, challenges, quot_weights)                 # [Q,3]
    quotient_poly = xcoset_interpolate(quotient_codeword, off)                       # stark.rs:1224-1231
    seg_polys = [quotient_poly[s::NUM_QUOTIENT_SEGMENTS] for s in range(NUM_QUOTIENT_SEGMENTS)]   # 1252-1263
    quot_rand = np.ascontiguousarray(quot_rand, dtype=np.uint64).reshape(-1, 3)
    assert quot_rand.shape[0] == d["num_quotient_randomizer_coefficients"]
    # randomize_quotient_segments (stark.rs:1302-1356)
    polys = [np.array(p, dtype=np.uint64) for p in seg_polys] + [quot_rand]
    zeta_k = pow(ZETA, NUM_QUOTIENT_SEGMENTS, P)
    for i in range(NUM_QUOTIENT_SEGMENTS - 1, -1, -1):
        nxt = polys[i + 1]
        scale = (-pow(ZETA, i, P)) % P
        addend = np.zeros_like(nxt)
        acc = 1
        for j in range(nxt.shape[0]):                                # scale(zeta^k): f(X) -> f(zeta^k X)
            f = scale * acc % P
            addend[j] = [int(nxt[j, 0]) * f % P, int(nxt[j, 1]) * f % P, int(nxt[j, 2]) * f % P]
            acc = acc * zeta_k % P
        m = max(polys[i].shape[0], addend.shape[0])
        s = np.zeros((m, 3), dtype=np.uint64)
        for j in range(m):
            a = as_x(polys[i][j:j + 1])[0] if j < polys[i].shape[0] else F.X_ZERO
            b = as_x(addend[j:j + 1])[0] if j < addend.shape[0] else F.X_ZERO
            s[j] = F.xadd(a, b)
        polys[i] = s
    seg_codewords = [xcoset_evaluate(p, off, N) for p in polys]                      # 5 x [N,3]
    quot_nodes = corc.merkle_build(xrow_digests(seg_codewords))
    ps.enqueue("MerkleRoot", [int(v) for v in quot_nodes[1]])

    # out-of-domain rows
    alpha = ps.sample_scalars(1)[0]
    omega = F.primitive_root_of_unity(n)
    alpha_next = F.xscale(alpha, omega)
    main_x = [[(int(v), 0, 0) for v in col] for col in main_trace]
    main_rand_x = [[(int(v), 0, 0) for v in col] for col in main_rand]
    aux_x = [as_x(aux_trace[q]) for q in range(NUM_AUX_COLUMNS)]
    aux_rand_x = [as_x(aux_rand[q]) for q in range(NUM_AUX_COLUMNS)]
    ood_main = out_of_domain_row(main_x, main_rand_x, alpha)
    ps.enqueue("OutOfDomainMainRow", ood_main)
    ood_aux = out_of_domain_row(aux_x, aux_rand_x, alpha)
    ps.enqueue("OutOfDomainAuxRow", ood_aux)
    ood_main_next = out_of_domain_row(main_x, main_rand_x, alpha_next)
    ps.enqueue("OutOfDomainMainRow", ood_main_next)
    ood_aux_next = out_of_domain_row(aux_x, aux_rand_x, alpha_next)
    ps.enqueue("OutOfDomainAuxRow", ood_aux_next)
    alpha_pow = F.xpow(alpha, NUM_QUOTIENT_SEGMENTS)
    alpha_zeta_pow = F.xpow(F.xscale(alpha, ZETA), NUM_QUOTIENT_SEGMENTS)
    ood_p = [xpoly_eval(p, alpha_pow) for p in polys[:-1]]
    ps.enqueue("OutOfDomainQuotientSegments", ood_p)
    ood_r = [xpoly_eval(p, alpha_zeta_pow) for p in polys[1:]]
    ps.enqueue("OutOfDomainQuotientSegments", ood_r)

    # combination codeword
    wm, wq, wd = ps.sample_scalars(3)                                                # stark.rs:2166-2209
    w_main_aux = xpows(wm, NUM_MAIN_COLUMNS + NUM_AUX_COLUMNS)
    w_quot = xpows(wq, NUM_RANDOMIZED_QUOTIENT_SEGMENTS)
    w_deep = xpows(wd, NUM_DEEP_CODEWORD_COMPONENTS)
    # weighted_sum_of_columns (master_table.rs:512-542), as the weighted sum of the column interpolants
    main_coef = column_interpolants(main_trace, main_rand)                           # [379, 2n]
    aux_coef = column_interpolants(aux_planar, aux_rand_planar)                      # [273, 2n]
    comb = [F.X_ZERO] * (2 * n)
    for c in range(NUM_MAIN_COLUMNS):
        w = w_main_aux[c]
        col = main_coef[c]
        for j in range(2 * n):
            v = int(col[j])
            if v:
                comb[j] = F.xadd(comb[j], F.xscale(w, v))
    for q in range(NUM_AUX_COLUMNS):
        w = w_main_aux[NUM_MAIN_COLUMNS + q]
        c0, c1, c2 = aux_coef[3 * q], aux_coef[3 * q + 1], aux_coef[3 * q + 2]
        for j in range(2 * n):
            v = (int(c0[j]), int(c1[j]), int(c2[j]))
            if v != F.X_ZERO:
                comb[j] = F.xadd(comb[j], F.xmul(w, v))
    main_aux_codeword = xcoset_evaluate(comb, off, N)

    def wsum(ps_, ws):
        m = max(p.shape[0] for p in ps_)
        out = [F.X_ZERO] * m
        for p, w in zip(ps_, ws):
            for j in range(p.shape[0]):
                out[j] = F.xadd(out[j], F.xmul(w, tuple(int(v) for v in p[j])))
        return out

    shared = wsum(polys[1:-1], w_quot[1:-1])
    poly_p = wsum([polys[0], np.array(shared, dtype=np.uint64)], [w_quot[0], F.X_ONE])
    poly_r = wsum([polys[-1], np.array(shared, dtype=np.uint64)], [w_quot[-1], F.X_ONE])
    cw_p = xcoset_evaluate(poly_p, off, N)
    cw_r = xcoset_evaluate(poly_r, off, N)

    xs = domain_values(off, N)

    def deep(cw, point, value):                                                      # stark.rs:1360-1379, 2096-2103
        inv = F.xbatch_inversion([F.xsub((x, 0, 0), point) for x in xs])
        return [F.xmul(F.xsub(tuple(int(v) for v in cw[i]), value), inv[i]) for i in range(N)]

    comps = [deep(main_aux_codeword, alpha, xpoly_eval(comb, alpha)),
             deep(main_aux_codeword, alpha_next, xpoly_eval(comb, alpha_next)),
             deep(cw_p, alpha_pow, xpoly_eval(poly_p, alpha_pow)),
             deep(cw_r, alpha_zeta_pow, xpoly_eval(poly_r, alpha_zeta_pow))]
    combination = np.zeros((N, 3), dtype=np.uint64)
    for i in range(N):
        acc = F.X_ZERO
        for k in range(NUM_DEEP_CODEWORD_COMPONENTS):
            acc = F.xadd(acc, F.xmul(comps[k][i], w_deep[k]))
        combination[i] = acc

    if d["ldt"] == "stir":
        revealed, fri_rounds = stir_mod.prove(ps, [tuple(int(t) for t in v) for v in combination], d["stir"]), []
    else:
        revealed, fri_rounds = fri_prove(ps, combination, d)

    # zero-knowledge guard (stark.rs:648-663)
    if alpha_pow[1] == 0 and alpha_pow[2] == 0:
        pts = {xs[i] for i in revealed}
        if alpha_pow[0] in pts or alpha_pow[0] * pow(ZETA, NUM_QUOTIENT_SEGMENTS, P) % P in pts:
            raise ValueError("ZeroKnowledgeViolation")

    # open rows
    def auth(nodes):
        return [[int(v) for v in nodes[k]] for k in merkle.auth_structure_node_indices(N, revealed)]

    ps.enqueue("MasterMainTableRows", [[int(v) for v in main_lde[:, i * es]] for i in revealed])
    ps.enqueue("AuthenticationStructure", auth(main_nodes))
    ps.enqueue("MasterAuxTableRows", [[tuple(int(v) for v in aux_lde[3 * q:3 * q + 3, i * es]) for q in range(NUM_AUX_COLUMNS)]
                                      for i in revealed])
    ps.enqueue("AuthenticationStructure", auth(aux_nodes))
    ps.enqueue("QuotientSegmentsElements", [[tuple(int(v) for v in seg_codewords[s][i]) for s in range(5)] for i in revealed])
    ps.enqueue("AuthenticationStructure", auth(quot_nodes))

    if keep:
        art.update(main_lde=main_lde, aux_lde=aux_lde, main_root=main_nodes[1], aux_root=aux_nodes[1],
                   quot_root=quot_nodes[1], challenges=challenges, quot_weights=quot_weights,
                   quotient_codeword=quotient_codeword, segment_codewords=seg_codewords, segment_polys=polys,
                   alpha=alpha, ood=(ood_main, ood_aux, ood_main_next, ood_aux_next, ood_p, ood_r),
                   combination=combination, fri_roots=[r[2][1] for r in fri_rounds], revealed=revealed)
    return ps.encode(), art


# ---- verifier ---------------------------------------------------------------------------------------
def verify(stark, claim, proof_words, check_air=True):
    """Verifier::verify (stark.rs:1388-1763).  `check_air=False` skips only the out-of-domain
    AIR/quotient identity (1469-1540), for proofs over synthetic (non-satisfying) traces."""
    ps = codec.decode_proof(proof_words)
    ps.alter_fiat_shamir_state_with(claim.encode())
    log2_ph = ps.dequeue("Log2PaddedHeight")
    if log2_ph >= 32:
        raise ValueError("Log2PaddedHeightTooLarge")
    d = stark.derive(1 << log2_ph)
    N, n = d["ldt_len"], d["trace_len"]
    height = N.bit_length() - 1
    main_root = ps.dequeue("MerkleRoot")
    challenges = derive_challenges(ps.sample_scalars(NUM_SAMPLED_CHALLENGES), claim)
    aux_root = ps.dequeue("MerkleRoot")
    w0 = ps.sample_scalars(1)[0]
    quot_root = ps.dequeue("MerkleRoot")
    omega = F.primitive_root_of_unity(n)
    alpha = ps.sample_scalars(1)[0]
    alpha_next = F.xscale(alpha, omega)
    alpha_pow = F.xpow(alpha, NUM_QUOTIENT_SEGMENTS)
    alpha_zeta_pow = F.xpow(F.xscale(alpha, ZETA), NUM_QUOTIENT_SEGMENTS)
    ood_main = ps.dequeue("OutOfDomainMainRow")
    ood_aux = ps.dequeue("OutOfDomainAuxRow")
    ood_main_next = ps.dequeue("OutOfDomainMainRow")
    ood_aux_next = ps.dequeue("OutOfDomainAuxRow")
    ood_p = ps.dequeue("OutOfDomainQuotientSegments")
    ood_r = ps.dequeue("OutOfDomainQuotientSegments")

    if check_air:
        from airgen.evaluate import evaluate_constraints
        a = air()
        quot_weights = xpows(w0, sum(len(a.constraints[c]) for c in ("init", "cons", "tran", "term")))
        zi = F.xinv(F.xsub(alpha, F.X_ONE))
        zc = F.xinv(F.xsub(F.xpow(alpha, n), F.X_ONE))
        except_last = F.xsub(alpha, (F.inv(omega), 0, 0))
        zinv = {"init": zi, "cons": zc, "tran": F.xmul(except_last, zc), "term": F.xinv(except_last)}
        total, k = F.X_ZERO, 0
        for cat in ("init", "cons", "tran", "term"):
            for v in evaluate_constraints(a.constraints[cat], ood_main, ood_aux, ood_main_next, ood_aux_next, challenges):
                total = F.xadd(total, F.xmul(quot_weights[k], F.xmul(v, zinv[cat])))
                k += 1
        lhs = F.X_ZERO
        for i in range(NUM_QUOTIENT_SEGMENTS):
            lhs = F.xadd(lhs, F.xmul(F.xpow(alpha, i), ood_p[i]))
        az = F.xscale(alpha, ZETA)
        for i in range(NUM_QUOTIENT_SEGMENTS):
            lhs = F.xadd(lhs, F.xmul(F.xpow(az, i), ood_r[i]))
        if total != lhs:
            raise ValueError("OutOfDomainQuotientValueMismatch")

    wm, wq, wd = ps.sample_scalars(3)
    w_main_aux = xpows(wm, NUM_MAIN_COLUMNS + NUM_AUX_COLUMNS)
    w_quot = xpows(wq, NUM_RANDOMIZED_QUOTIENT_SEGMENTS)
    w_deep = xpows(wd, NUM_DEEP_CODEWORD_COMPONENTS)

    def lin(main_row, aux_row):
        acc = F.X_ZERO
        for w, v in zip(w_main_aux, list(main_row) + list(aux_row)):
            acc = F.xadd(acc, F.xmul(w, v if isinstance(v, tuple) else (int(v), 0, 0)))
        return acc

    ood_curr_value = lin(ood_main, ood_aux)
    ood_next_value = lin(ood_main_next, ood_aux_next)
    ood_p_value, ood_r_value = F.X_ZERO, F.X_ZERO
    for e, w in zip(ood_p, w_quot[:-1]): ood_p_value = F.xadd(ood_p_value, F.xmul(e, w))
    for e, w in zip(ood_r, w_quot[1:]): ood_r_value = F.xadd(ood_r_value, F.xmul(e, w))

    indices, ldt_values = stir_mod.verify(ps, d["stir"]) if d["ldt"] == "stir" else fri_verify(ps, d)
    q = d["num_first_round_queries"]                            # ldt.num_first_round_queries(), stark.rs:1581-1586
    if len(indices) != q or len(ldt_values) != q:
        raise ValueError("IncorrectNumberOfRowIndices")

    main_rows = ps.dequeue("MasterMainTableRows")
    main_auth = ps.dequeue("AuthenticationStructure")
    if len(main_rows) != q: raise ValueError("IncorrectNumberOfMainTableRows")
    leafs = [tip5.hash_varlen(r) for r in main_rows]
    if not merkle.verify_inclusion(main_root, height, list(zip(indices, leafs)), main_auth):
        raise ValueError("MainCodewordAuthenticationFailure")
    aux_rows = ps.dequeue("MasterAuxTableRows")
    aux_auth = ps.dequeue("AuthenticationStructure")
    if len(aux_rows) != q: raise ValueError("IncorrectNumberOfAuxTableRows")
    leafs = [tip5.hash_varlen(codec.enc_xfes(r)) for r in aux_rows]
    if not merkle.verify_inclusion(aux_root, height, list(zip(indices, leafs)), aux_auth):
        raise ValueError("AuxiliaryCodewordAuthenticationFailure")
    quot_rows = ps.dequeue("QuotientSegmentsElements")
    quot_auth = ps.dequeue("AuthenticationStructure")
    if len(quot_rows) != q: raise ValueError("IncorrectNumberOfQuotientSegmentElements")
    leafs = [tip5.hash_varlen(codec.enc_xfes(r)) for r in quot_rows]
    if not merkle.verify_inclusion(quot_root, height, list(zip(indices, leafs)), quot_auth):
        raise ValueError("QuotientCodewordAuthenticationFailure")

    g = F.primitive_root_of_unity(N)

    def deep_update(x, value, point, ood_value):
        return F.xmul(F.xsub(value, ood_value), F.xinv(F.xsub((x, 0, 0), point)))

    for idx, mrow, arow, qrow, revealed_value in zip(indices, main_rows, aux_rows, quot_rows, ldt_values):
        x = d["ldt_offset"] * pow(g, idx, P) % P
        ma = lin(mrow, arow)
        shared = F.X_ZERO
        for e, w in zip(qrow[1:-1], w_quot[1:-1]): shared = F.xadd(shared, F.xmul(e, w))
        for_p = F.xadd(F.xmul(w_quot[0], qrow[0]), shared)
        for_r = F.xadd(F.xmul(w_quot[-1], qrow[-1]), shared)
        comps = [deep_update(x, ma, alpha, ood_curr_value), deep_update(x, ma, alpha_next, ood_next_value),
                 deep_update(x, for_p, alpha_pow, ood_p_value), deep_update(x, for_r, alpha_zeta_pow, ood_r_value)]
        acc = F.X_ZERO
        for c, w in zip(comps, w_deep): acc = F.xadd(acc, F.xmul(c, w))
        if tuple(revealed_value) != acc:
            raise ValueError("CombinationCodewordMismatch")
    if ps.index != len(ps.items):
        raise ValueError("SuperfluousProofItems")
    return True

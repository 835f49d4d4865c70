"""BFieldCodec encodings of the proof items + the Fiat-Shamir proof stream.

Restates `triton-vm/src/proof_stream.rs:19-125`, `proof_item.rs:96-147` (variant order =
discriminant, which items are absorbed into the sponge) and `proof.rs:37-88`; the encoding
rules themselves are twenty-first 2.0's `BFieldCodec`.  They were recalled, and are now PINNED by the
reference's two whole-proof known-answer digests (proof.rs:200-226, stark.rs:2433-2460; tests/test_golden.py):

  * BFieldElement -> 1 word; XFieldElement -> 3 words (c0,c1,c2); Digest -> 5; u32 -> 1;
    [T; N] of statically sized T -> N*len(T) words, no prefix;
  * Vec<T>: element count, then the elements; dynamically sized elements are each prefixed
    with their length;
  * derived struct: every dynamically sized field is prefixed with its encoded length;
    fields are emitted in REVERSE declaration order (STRUCT_FIELDS_REVERSED);
  * derived enum: variant index, then the payload field (length-prefixed when dynamic);
  * Polynomial<XFE>: a one-field struct — the coefficient Vec (trailing zeros stripped) behind its encoded length.

TEST INFRASTRUCTURE ONLY."""
from . import tip5


STRUCT_FIELDS_REVERSED = True
# Polynomial<FF> { coefficients } encoded like a derived one-field struct: the coefficient Vec's encoding prefixed with
# its length (confirmed by the whole-proof digests).  False: the bare Vec encoding (kept for the search tool in tests/golden).
POLYNOMIAL_AS_STRUCT = True

# proof_item.rs:96-147 — (variant index, absorbed into Fiat-Shamir?)
ITEMS = {
    "MerkleRoot": (0, True), "Log2PaddedHeight": (1, True), "OutOfDomainMainRow": (2, True),
    "OutOfDomainAuxRow": (3, True), "OutOfDomainQuotientSegments": (4, True), "Polynomial": (5, True),
    "StirOutOfDomainValues": (6, True), "AuthenticationStructure": (7, False), "MasterMainTableRows": (8, False),
    "MasterAuxTableRows": (9, False), "QuotientSegmentsElements": (10, False), "FriCodeword": (11, False),
    "FriResponse": (12, False), "StirResponse": (13, False),
}
STATIC_PAYLOAD = {"MerkleRoot", "Log2PaddedHeight", "OutOfDomainMainRow", "OutOfDomainAuxRow", "OutOfDomainQuotientSegments"}


def enc_xfes(xs):
    out = []
    for x in xs:
        out += [int(x[0]), int(x[1]), int(x[2])]
    return out


def enc_vec_xfe(xs):
    return [len(xs)] + enc_xfes(xs)


def enc_vec_vec_xfe(vs):
    """Vec<Vec<XFE>>: count, then each (dynamically sized) element prefixed with its encoded length"""
    out = [len(vs)]
    for v in vs:
        e = enc_vec_xfe(v)
        out += [len(e)] + e
    return out


def enc_vec_digest(ds):
    out = [len(ds)]
    for d in ds:
        out += [int(v) for v in d]
    return out


def enc_struct(fields):
    """fields: list of (encoding, is_dynamic) in declaration order"""
    out = []
    for enc, dyn in (reversed(fields) if STRUCT_FIELDS_REVERSED else fields):
        if dyn:
            out.append(len(enc))
        out += enc
    return out


def enc_polynomial_xfe(coeffs):
    c = list(coeffs)
    while c and tuple(c[-1]) == (0, 0, 0):
        c.pop()
    e = enc_vec_xfe(c)
    return [len(e)] + e if POLYNOMIAL_AS_STRUCT else e


def encode_payload(kind, payload):
    if kind == "MerkleRoot": return [int(v) for v in payload]
    if kind == "Log2PaddedHeight": return [int(payload)]
    if kind in ("OutOfDomainMainRow", "OutOfDomainAuxRow", "OutOfDomainQuotientSegments"): return enc_xfes(payload)
    if kind == "Polynomial": return enc_polynomial_xfe(payload)
    if kind in ("StirOutOfDomainValues", "FriCodeword"): return enc_vec_xfe(payload)
    if kind == "AuthenticationStructure": return enc_vec_digest(payload)
    if kind == "MasterMainTableRows":          # Vec<[BFE; 379]>
        out = [len(payload)]
        for row in payload: out += [int(v) for v in row]
        return out
    if kind in ("MasterAuxTableRows", "QuotientSegmentsElements"):   # Vec<[XFE; N]>
        out = [len(payload)]
        for row in payload: out += enc_xfes(row)
        return out
    if kind == "FriResponse":                  # fri.rs:99-107 {queried_leaves, auth_structure}
        leaves, auth = payload
        return enc_struct([(enc_vec_xfe(leaves), True), (enc_vec_digest(auth), True)])
    if kind == "StirResponse":                 # stir.rs:149-158 {queried_leafs: Vec<Vec<XFE>>, auth_structure}
        leafs, auth = payload
        return enc_struct([(enc_vec_vec_xfe(leafs), True), (enc_vec_digest(auth), True)])
    raise ValueError(kind)


def encode_item(kind, payload):
    idx, _ = ITEMS[kind]
    enc = encode_payload(kind, payload)
    if kind in STATIC_PAYLOAD:
        return [idx] + enc
    return [idx, len(enc)] + enc


def encode_claim(program_digest, version, inp, out):
    """proof.rs:68-88: Claim {program_digest: Digest, version: u32, input: Vec<BFE>, output: Vec<BFE>}"""
    return enc_struct([([int(v) for v in program_digest], False), ([int(version)], False),
                       ([len(inp)] + [int(v) for v in inp], True), ([len(out)] + [int(v) for v in out], True)])


class ProofStream:
    def __init__(self):
        self.items = []            # (kind, payload)
        self.sponge = tip5.Tip5()  # Tip5::init(): variable-length domain (proof_stream.rs:24)
        self.index = 0

    def alter_fiat_shamir_state_with(self, encoding):
        self.sponge.pad_and_absorb_all(encoding)

    def enqueue(self, kind, payload):
        if ITEMS[kind][1]:
            self.alter_fiat_shamir_state_with(encode_item(kind, payload))
        self.items.append((kind, payload))

    def dequeue(self, expect):
        if self.index >= len(self.items):
            raise ValueError("EmptyQueue")
        kind, payload = self.items[self.index]
        if kind != expect:
            raise ValueError(f"UnexpectedItem: expected {expect}, got {kind}")
        if ITEMS[kind][1]:
            self.alter_fiat_shamir_state_with(encode_item(kind, payload))
        self.index += 1
        return payload

    def sample_scalars(self, n): return self.sponge.sample_scalars(n)
    def sample_indices(self, upper_bound, n): return self.sponge.sample_indices(upper_bound, n)

    def encode(self):
        """Proof(Vec<BFE>) = ProofStream{items}.encode()  (proof_stream.rs:8-17, 110-119)"""
        body = [len(self.items)]
        for kind, payload in self.items:
            e = encode_item(kind, payload)
            body += [len(e)] + e
        return [len(body)] + body


# ---- decoding (verifier side) -----------------------------------------------------------------
class _Reader:
    def __init__(self, words, pos=0, end=None):
        self.w, self.pos, self.end = words, pos, len(words) if end is None else end

    def take(self, n=1):
        if self.pos + n > self.end:
            raise ValueError("SequenceTooShort")
        out = self.w[self.pos:self.pos + n]
        self.pos += n
        return out

    def one(self): return self.take(1)[0]
    def xfe(self): return tuple(self.take(3))
    def digest(self): return list(self.take(5))
    def vec_xfe(self): return [self.xfe() for _ in range(self.one())]
    def vec_digest(self): return [self.digest() for _ in range(self.one())]
    def done(self): return self.pos == self.end


def decode_payload(kind, words, num_main=379, num_aux=91, num_seg=5):
    r = _Reader(words)
    if kind == "MerkleRoot": out = r.digest()
    elif kind == "Log2PaddedHeight": out = r.one()
    elif kind == "OutOfDomainMainRow": out = [r.xfe() for _ in range(num_main)]
    elif kind == "OutOfDomainAuxRow": out = [r.xfe() for _ in range(num_aux)]
    elif kind == "OutOfDomainQuotientSegments": out = [r.xfe() for _ in range(num_seg - 1)]
    elif kind in ("Polynomial", "StirOutOfDomainValues", "FriCodeword"):
        if kind == "Polynomial" and POLYNOMIAL_AS_STRUCT:
            if r.one() != len(words) - 1: raise ValueError("field length mismatch")
        out = r.vec_xfe()
        if kind == "Polynomial" and out and out[-1] == (0, 0, 0):
            raise ValueError("TrailingZerosInPolynomialEncoding")
    elif kind == "AuthenticationStructure": out = r.vec_digest()
    elif kind == "MasterMainTableRows": out = [list(r.take(num_main)) for _ in range(r.one())]
    elif kind == "MasterAuxTableRows": out = [[r.xfe() for _ in range(num_aux)] for _ in range(r.one())]
    elif kind == "QuotientSegmentsElements": out = [[r.xfe() for _ in range(num_seg)] for _ in range(r.one())]
    elif kind == "FriResponse":
        names = ["leaves", "auth"]
        order = list(reversed(names)) if STRUCT_FIELDS_REVERSED else names
        got = {}
        for nm in order:
            ln = r.one()
            sub = _Reader(r.take(ln))
            got[nm] = sub.vec_xfe() if nm == "leaves" else sub.vec_digest()
            if not sub.done(): raise ValueError("field length mismatch")
        out = (got["leaves"], got["auth"])
    elif kind == "StirResponse":
        names = ["leafs", "auth"]
        order = list(reversed(names)) if STRUCT_FIELDS_REVERSED else names
        got = {}
        for nm in order:
            ln = r.one()
            sub = _Reader(r.take(ln))
            if nm == "leafs":
                vs = []
                for _ in range(sub.one()):
                    inner = _Reader(sub.take(sub.one()))
                    vs.append(inner.vec_xfe())
                    if not inner.done(): raise ValueError("inner length mismatch")
                got[nm] = vs
            else:
                got[nm] = sub.vec_digest()
            if not sub.done(): raise ValueError("field length mismatch")
        out = (got["leafs"], got["auth"])
    else:
        raise ValueError(kind)
    if not r.done():
        raise ValueError("payload length mismatch")
    return out


def decode_proof(words):
    by_idx = {v[0]: k for k, v in ITEMS.items()}
    r = _Reader([int(w) for w in words])
    ln = r.one()
    if ln != len(words) - 1: raise ValueError("proof length prefix mismatch")
    ps = ProofStream()
    for _ in range(r.one()):
        item_len = r.one()
        item = r.take(item_len)
        kind = by_idx[item[0]]
        body = item[1:] if kind in STATIC_PAYLOAD else item[2:]
        if kind not in STATIC_PAYLOAD and item[1] != len(body): raise ValueError("payload prefix mismatch")
        ps.items.append((kind, decode_payload(kind, body)))
    if not r.done(): raise ValueError("trailing words")
    return ps

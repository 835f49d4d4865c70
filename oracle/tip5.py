"""Tip5 permutation + sponge, pure Python ints (slow; small inputs / KATs only).

Follows `tips/tip-0005/tip-0005.md:21-81` (parameters, S-boxes, MDS first
column, round-constant derivation, sponge modes) and the call-site semantics
in the reference: absorb overwrites the rate then permutes
(`triton-vm/src/aet.rs:190-196`, `table/master_table.rs:696-699`); squeeze
returns lanes 0..9 then permutes (`vm.rs:730-738`); `hash_varlen` semantics
(`master_table.rs:703-715`).  `sample_scalars` / `sample_indices` restate
twenty-first 2.0 (`proof_stream.rs:93-102` are the call sites).

State is kept in canonical form; the split-and-lookup S-box converts to the
Montgomery representation and back, exactly as tip-0005.md:52-61 defines it.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import hashlib
from .field import P, R, R_INV

STATE = 16
RATE = 10
CAPACITY = 6
DIGEST = 5
ROUNDS = 5
NUM_SPLIT_AND_LOOKUP = 4

LOOKUP_TABLE = [((x + 1) ** 3 - 1) % 257 for x in range(256)]
assert all(v < 256 for v in LOOKUP_TABLE)

_sha = hashlib.sha256(b"Tip5").digest()
MDS_FIRST_COLUMN = [int.from_bytes(_sha[2 * i:2 * i + 2], "little") for i in range(16)]


def _round_constants():
    import blake3
    out = []
    for i in range(STATE * ROUNDS):
        d = blake3.blake3(b"Tip5" + bytes([i])).digest()
        v = int.from_bytes(d[:16], "little") % P
        out.append(v * R_INV % P)
    return out

try:
    ROUND_CONSTANTS = _round_constants()
except ImportError:  # blake3 module absent: fall back to the committed table
    from .tip5_constants import ROUND_CONSTANTS  # type: ignore


def _split_and_lookup(x):
    m = x * R % P
    b = m.to_bytes(8, "little")
    m2 = int.from_bytes(bytes(LOOKUP_TABLE[v] for v in b), "little")
    return m2 * R_INV % P


def permutation(state):
    s = list(state)
    for rnd in range(ROUNDS):
        for i in range(NUM_SPLIT_AND_LOOKUP):
            s[i] = _split_and_lookup(s[i])
        for i in range(NUM_SPLIT_AND_LOOKUP, STATE):
            s[i] = pow(s[i], 7, P)
        # circulant MDS: entry (r, c) = first_column[(r - c) mod 16]  (triton-air/src/table/hash.rs:49-56)
        t = [sum(MDS_FIRST_COLUMN[(r - c) % 16] * s[c] for c in range(16)) % P for r in range(16)]
        s = [(t[i] + ROUND_CONSTANTS[16 * rnd + i]) % P for i in range(16)]
    return s


class Tip5:
    """Sponge object; `Tip5()` == `Tip5::init()` == variable-length domain."""
    def __init__(self, fixed_length=False):
        self.state = [0] * STATE
        if fixed_length:
            for i in range(RATE, STATE):
                self.state[i] = 1

    def absorb(self, chunk):
        assert len(chunk) == RATE
        self.state[:RATE] = [c % P for c in chunk]
        self.state = permutation(self.state)

    def pad_and_absorb_all(self, words):
        words = list(words)
        n_full = len(words) // RATE
        for i in range(n_full):
            self.absorb(words[RATE * i:RATE * (i + 1)])
        rem = words[RATE * n_full:]
        last = rem + [1] + [0] * (RATE - len(rem) - 1)
        self.absorb(last)

    def squeeze(self):
        out = self.state[:RATE]
        self.state = permutation(self.state)
        return out

    def sample_scalars(self, n):
        num_squeezes = (3 * n + RATE - 1) // RATE
        els = []
        for _ in range(num_squeezes):
            els += self.squeeze()
        return [tuple(els[3 * i:3 * i + 3]) for i in range(n)]

    def sample_indices(self, upper_bound, n):
        assert upper_bound & (upper_bound - 1) == 0
        out, buf = [], []
        while len(out) != n:
            if not buf:
                buf = list(reversed(self.squeeze()))
            e = buf.pop()
            if e != P - 1:
                out.append((e & 0xFFFFFFFF) % upper_bound)
        return out


def hash_10(inp):
    s = Tip5(fixed_length=True)
    s.state[:RATE] = [c % P for c in inp]
    return permutation(s.state)[:DIGEST]

def hash_pair(l, r):
    return hash_10(list(l) + list(r))

def hash_varlen(words):
    s = Tip5()
    s.pad_and_absorb_all(words)
    return s.state[:DIGEST]

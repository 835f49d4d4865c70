"""Restatement of the slice of `rand = "0.10.1"` (reference Cargo.toml:96) and of twenty-first's
`Distribution<BFieldElement/XFieldElement>` that the reference's seeded tests use
(`StdRng::seed_from_u64`, `rng.random()`); only needed to reach the AIR known-answer test
`air_constraints_evaluators_have_not_changed` (master_table.rs:2327-2415).

Recalled from the published crates (not in-tree); pinned by the AIR known-answer test and by the two whole-proof digests
(tests/test_golden.py), which consume ~10^6 draws:
  * StdRng = ChaCha12, 64-word output buffer (4 blocks), next_u64 = two consecutive words (lo, hi);
  * seed_from_u64 expands the u64 with PCG32 (MUL 6364136223846793005, INC 11634580027462260723);
  * BFieldElement sample = BFieldElement::new(rng.random_range(0..=BFieldElement::MAX)) with
    rand's single-sample "Canon" widening-multiply method; XFieldElement = 3 such samples.
TEST INFRASTRUCTURE ONLY."""
import struct

MASK32 = 0xFFFFFFFF
MASK64 = (1 << 64) - 1
P = (1 << 64) - (1 << 32) + 1


def _rotl(x, n): return ((x << n) & MASK32) | (x >> (32 - n))


def chacha_block(key_words, counter, stream, rounds):
    const = [0x61707865, 0x3320646e, 0x79622d32, 0x6b206574]
    s = const + list(key_words) + [counter & MASK32, (counter >> 32) & MASK32, stream & MASK32, (stream >> 32) & MASK32]
    w = list(s)

    def qr(a, b, c, d):
        w[a] = (w[a] + w[b]) & MASK32; w[d] = _rotl(w[d] ^ w[a], 16)
        w[c] = (w[c] + w[d]) & MASK32; w[b] = _rotl(w[b] ^ w[c], 12)
        w[a] = (w[a] + w[b]) & MASK32; w[d] = _rotl(w[d] ^ w[a], 8)
        w[c] = (w[c] + w[d]) & MASK32; w[b] = _rotl(w[b] ^ w[c], 7)

    for _ in range(rounds // 2):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(a + b) & MASK32 for a, b in zip(w, s)]


class StdRng:
    def __init__(self, seed32, rounds=12):
        assert len(seed32) == 32
        self.key = struct.unpack("<8I", bytes(seed32))
        self.counter = 0
        self.rounds = rounds
        self.buf = []
        self.index = 64

    @classmethod
    def seed_from_u64(cls, state, rounds=12):
        MUL, INC = 6364136223846793005, 11634580027462260723
        seed = b""
        for _ in range(8):
            state = (state * MUL + INC) & MASK64
            xorshifted = (((state >> 18) ^ state) >> 27) & MASK32
            rot = state >> 59
            x = ((xorshifted >> rot) | (xorshifted << ((32 - rot) & 31))) & MASK32
            seed += struct.pack("<I", x)
        return cls(seed, rounds)

    def _generate(self):
        self.buf = []
        for _ in range(4):
            self.buf += chacha_block(self.key, self.counter, 0, self.rounds)
            self.counter += 1
        self.index = 0

    def next_u32(self):
        if self.index >= 64:
            self._generate()
        v = self.buf[self.index]
        self.index += 1
        return v

    def next_u64(self):
        if self.index < 63:
            lo, hi = self.buf[self.index], self.buf[self.index + 1]
            self.index += 2
        elif self.index == 63:
            lo = self.buf[63]
            self._generate()
            hi = self.buf[0]
            self.index = 1
        else:
            self._generate()
            lo, hi = self.buf[0], self.buf[1]
            self.index = 2
        return lo | (hi << 32)

    def random_range_inclusive_u64(self, low, high):
        rng_range = (high - low + 1) & MASK64
        if rng_range == 0:
            return self.next_u64()
        prod = self.next_u64() * rng_range
        result, lo_order = prod >> 64, prod & MASK64
        if lo_order > ((-rng_range) & MASK64):
            new_hi_order = (self.next_u64() * rng_range) >> 64
            if lo_order + new_hi_order > MASK64:
                result += 1
        return (low + result) & MASK64

    def bfe(self):
        return self.random_range_inclusive_u64(0, P - 1) % P

    def xfe(self):
        return (self.bfe(), self.bfe(), self.bfe())

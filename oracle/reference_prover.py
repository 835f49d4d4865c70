"""`Prover::prove` driven exactly like the reference drives it — from a program, its inputs and ONE 32-byte randomness
seed (stark.rs:331-372, 1316-1322; master_table.rs:423-434, 630-662, 1006-1025) — on top of the oracle's VM, table fill,
AIR-derived extension and prover.  With it the reference's whole-proof known-answer tests are reachable; they pin every
convention of the pipeline at once (tests/test_golden.py).

TEST INFRASTRUCTURE ONLY."""
import numpy as np

from . import corc, stark as S, tip5, tracegen as tg
from .rand_compat import StdRng

NUM_MAIN, NUM_AUX = 379, 91


def offset_rng_seed(seed, offset):
    """master_table.rs:630-662: little-endian addition of `offset` into the 32 seed bytes, carries propagated"""
    out, carry = bytearray(seed), 0
    addend = int(offset).to_bytes(8, "little")
    for i in range(32):
        s = out[i] + (addend[i] if i < 8 else 0) + carry
        out[i], carry = s & 0xFF, s >> 8
    return bytes(out)


def seed_from_rng(rng):
    """`rng.random::<[u8; 32]>()`: one `next_u32() as u8` per byte (rand's StandardUniform for arrays and u8)"""
    return bytes(rng.next_u32() & 0xFF for _ in range(32))


def randomness(seed, trace_len, num_trace_randomizers, num_quotient_randomizer_coefficients):
    """-> (main randomizer coefficients [379][h], aux ones [91][h][3], batch-randomizer column [n][3], quotient-segment
    randomizer [..][3]), each drawn from its own offset seed"""
    def bfes(sd, count):
        r = StdRng(sd)
        return [r.bfe() for _ in range(count)]

    def xfes(sd, count):
        r = StdRng(sd)
        return [r.xfe() for _ in range(count)]
    h = num_trace_randomizers
    mrand = np.array([bfes(offset_rng_seed(seed, i), h) for i in range(NUM_MAIN)], dtype=np.uint64)     # master_table.rs:429
    aux_seed = offset_rng_seed(seed, NUM_MAIN)                                                         # :1008-1009
    arand = np.array([xfes(offset_rng_seed(aux_seed, i), h) for i in range(NUM_AUX)], dtype=np.uint64)
    rcol = np.array(xfes(offset_rng_seed(aux_seed, NUM_AUX), trace_len), dtype=np.uint64)              # :1017-1025
    qrand = np.array(xfes(offset_rng_seed(seed, NUM_MAIN + NUM_AUX + 1), num_quotient_randomizer_coefficients),
                     dtype=np.uint64)                                                                  # stark.rs:1316-1321
    return mrand, arand, rcol, qrand


def instance(stark, words, public_input, seed, secret_input=(), initial_ram=None, secret_digests=()):
    """everything `Prover::prove` derives before the first commitment -> dict"""
    ex = tg.execute(words, public_input, secret_input, initial_ram, secret_digests)
    ph = tg.padded_height(words, public_input, secret_input, initial_ram, secret_digests)
    d = stark.derive(ph)
    n = d["trace_len"]
    T, digest, out = tg.main_table(words, public_input, n, secret_input, initial_ram, secret_digests)
    main = np.array(T, dtype=np.uint64)
    mrand, arand, rcol, qrand = randomness(seed, n, d["num_trace_randomizers"], d["num_quotient_randomizer_coefficients"])
    claim = S.Claim(digest, list(public_input), list(out))

    def extend(ch):
        return corc.aux_extend(main, np.asarray(ch, dtype=np.uint64).reshape(63, 3), rcol), arand
    return dict(stark=stark, claim=claim, main=main, main_rand=mrand, aux_rand=arand, randomizer_column=rcol, quot_rand=qrand,
                extend=extend, padded_height=ph, derived=d, execution=ex)


def prove(inst):
    proof, _ = S.prove(inst["stark"], inst["claim"], inst["main"], inst["main_rand"], inst["extend"], inst["quot_rand"],
                       padded_height=inst["padded_height"])
    return proof


def proof_digest(proof):
    """Tip5::hash(&proof): Proof(Vec<BFE>) is a one-field struct — length of the Vec encoding, count, words"""
    return [int(v) for v in tip5.hash_varlen([len(proof) + 1, len(proof)] + [int(v) for v in proof])]

"""GPU parity of the whole prove() path: the proof produced by libtvm_b200 must equal, word for
word, the proof of the CPU oracle (oracle/stark.py) on the same (claim, traces, randomizers), and
the oracle verifier must accept its Merkle / FRI / DEEP structure."""
import numpy as np
import pytest

from conftest import rand_bfes
from oracle import stark as S

pytestmark = pytest.mark.gpu


def synthetic_instance(security, log2_exp, padded_height, seed):
    st = S.Stark(security, log2_exp)
    d = st.derive(padded_height)
    rng = np.random.default_rng(seed)
    n, h = d["trace_len"], d["num_trace_randomizers"]
    main = rand_bfes(rng, (379, n))
    mrand = rand_bfes(rng, (379, h))
    qrand = rand_bfes(rng, (d["num_quotient_randomizer_coefficients"], 3))

    def aux_provider(challenges):
        # deterministic function of the challenges, like MasterMainTable::extend
        s = int(np.asarray(challenges, dtype=np.uint64).reshape(-1)[:8].sum() % (1 << 32))
        r = np.random.default_rng(seed * 1000003 + s)
        return rand_bfes(r, (91, n, 3)), rand_bfes(r, (91, h, 3))

    claim = S.Claim([11, 22, 33, 44, 55], [1, 2, 3], [4, 5])
    return st, d, claim, main, mrand, aux_provider, qrand


@pytest.mark.parametrize("security,log2_exp,padded_height,seed", [(4, 2, 16, 1), (8, 2, 64, 2), (32, 2, 256, 3),
                                                                  (4, 3, 16, 31), (8, 5, 32, 32),    # expansion 8, 32
                                                                  (4, 1, 16, 33)])                   # expansion 2
def test_proof_is_bit_exact_vs_oracle(backend, security, log2_exp, padded_height, seed):
    st, d, claim, main, mrand, aux_provider, qrand = synthetic_instance(security, log2_exp, padded_height, seed)
    want, _ = S.prove(st, claim, main, mrand, aux_provider, qrand, padded_height=padded_height)
    got = backend.prove((claim.program_digest, claim.input, claim.output), main, mrand, aux_provider, qrand,
                        security_level=security, log2_expansion=log2_exp, padded_height=padded_height)
    got = [int(v) for v in got]
    assert len(got) == len(want)
    assert got == want
    assert S.verify(st, claim, got, check_air=False)
    stages = dict(backend.last_prove_timings())
    assert "low-degree test" in stages and "quotient(AIR)" in stages


def test_proof_depends_on_every_input(backend):
    st, d, claim, main, mrand, aux_provider, qrand = synthetic_instance(4, 2, 16, 5)
    base = backend.prove((claim.program_digest, claim.input, claim.output), main, mrand, aux_provider, qrand,
                         security_level=4, log2_expansion=2, padded_height=16)
    main2 = main.copy(); main2[200, 3] ^= np.uint64(1)
    other = backend.prove((claim.program_digest, claim.input, claim.output), main2, mrand, aux_provider, qrand,
                          security_level=4, log2_expansion=2, padded_height=16)
    assert not np.array_equal(base, other)
    again = backend.prove((claim.program_digest, claim.input, claim.output), main, mrand, aux_provider, qrand,
                          security_level=4, log2_expansion=2, padded_height=16)
    assert np.array_equal(base, again)   # deterministic: all randomness is an input


def synthetic_stir_instance(security, padded_height, seed):
    st = S.Stark(security, 2, "stir")
    d = st.derive(padded_height)
    rng = np.random.default_rng(seed)
    n, h = d["trace_len"], d["num_trace_randomizers"]
    main, mrand = rand_bfes(rng, (379, n)), rand_bfes(rng, (379, h))
    qrand = rand_bfes(rng, (d["num_quotient_randomizer_coefficients"], 3))

    def aux_provider(challenges):
        s = int(np.asarray(challenges, dtype=np.uint64).reshape(-1)[:8].sum() % (1 << 32))
        r = np.random.default_rng(seed * 1000003 + s)
        return rand_bfes(r, (91, n, 3)), rand_bfes(r, (91, h, 3))

    return st, d, S.Claim([11, 22, 33, 44, 55], [1, 2, 3], [4, 5]), main, mrand, aux_provider, qrand


@pytest.mark.parametrize("security,padded_height,seed", [(8, 64, 11), (6, 256, 12), (10, 1024, 13)])
def test_stir_proof_is_bit_exact_vs_oracle(backend, security, padded_height, seed):
    """LdtChoice::Stir (what Stark::default() selects from padded height 2^16 on): 0, 2 and 3 full rounds."""
    import tvm_b200
    st, d, claim, main, mrand, aux_provider, qrand = synthetic_stir_instance(security, padded_height, seed)
    assert d["ldt"] == "stir"
    want, _ = S.prove(st, claim, main, mrand, aux_provider, qrand, padded_height=padded_height)
    got = backend.prove((claim.program_digest, claim.input, claim.output), main, mrand, aux_provider, qrand,
                        security_level=security, log2_expansion=2, padded_height=padded_height, ldt_choice=tvm_b200.LDT_STIR)
    got = [int(v) for v in got]
    assert len(got) == len(want)
    assert got == want
    assert S.verify(st, claim, got, check_air=False)


@pytest.mark.parametrize("log2_exp", [4, 1])
def test_stir_with_other_expansion_factors(backend, log2_exp):
    """expansion 16: the quotient domain is a sub-domain of the LDT-domain tables; expansion 2: the tables live on the
    (larger) quotient domain and the LDT domain is every second coset of it.  Cached and just-in-time tables."""
    import tvm_b200
    st = S.Stark(6, log2_exp, "stir")
    d = st.derive(64)
    rng = np.random.default_rng(41)
    n, h = d["trace_len"], d["num_trace_randomizers"]
    main, mrand = rand_bfes(rng, (379, n)), rand_bfes(rng, (379, h))
    qrand = rand_bfes(rng, (d["num_quotient_randomizer_coefficients"], 3))
    aux_t, aux_r = rand_bfes(rng, (91, n, 3)), rand_bfes(rng, (91, h, 3))
    claim = S.Claim([9, 8, 7, 6, 5], [1], [2, 3])
    want, _ = S.prove(st, claim, main, mrand, lambda ch: (aux_t, aux_r), qrand, padded_height=64)
    for mode in (2, 1):                           # cached tables, then just-in-time LDE
        backend.set_low_memory(mode)
        try:
            got = backend.prove((claim.program_digest, claim.input, claim.output), main, mrand, lambda ch: (aux_t, aux_r), qrand,
                                security_level=6, log2_expansion=log2_exp, padded_height=64, ldt_choice=tvm_b200.LDT_STIR)
        finally:
            backend.set_low_memory(0)
        assert [int(v) for v in got] == want, mode


@pytest.mark.parametrize("ldt", ["fri", "stir", "fri_tiles"])
def test_low_memory_mode_produces_the_same_proof(backend, ldt):
    """Just-in-time LDE (tables never stored, every coset re-evaluated for hashing, AIR and openings) vs cached tables.
    "fri_tiles": a trace domain of 2^12, where the square-tile NTT (ntt_tile.cu) evaluates one coset per call."""
    import tvm_b200
    if ldt == "fri_tiles":
        st, d, claim, main, mrand, aux_provider, qrand = synthetic_instance(4, 2, 4096, 23)
        kw = dict(security_level=4, log2_expansion=2, padded_height=4096)
    elif ldt == "stir":
        st, d, claim, main, mrand, aux_provider, qrand = synthetic_stir_instance(6, 256, 21)
        kw = dict(security_level=6, log2_expansion=2, padded_height=256, ldt_choice=tvm_b200.LDT_STIR)
    else:
        st, d, claim, main, mrand, aux_provider, qrand = synthetic_instance(32, 2, 256, 22)
        kw = dict(security_level=32, log2_expansion=2, padded_height=256)
    args = ((claim.program_digest, claim.input, claim.output), main, mrand, aux_provider, qrand)
    try:
        backend.set_low_memory(2)
        cached = backend.prove(*args, **kw)
        assert not backend.last_prove_low_memory
        backend.set_low_memory(1)
        jit = backend.prove(*args, **kw)
        assert backend.last_prove_low_memory
    finally:
        backend.set_low_memory(0)
    assert np.array_equal(cached, jit)


@pytest.mark.parametrize("ldt,padded_height", [("fri", 4), ("stir", 256)])
def test_default_security_on_small_instances(backend, ldt, padded_height):
    """Stark::default() parameters (security 160) on tiny padded heights: the trace domain is bumped by the number of
    trace randomizers (stark.rs:1885-1890); with STIR there is no full round at this size (final round only, 284
    queries).  Empty public input / output."""
    import tvm_b200
    st = S.Stark(160, 2, ldt)
    d = st.derive(padded_height)
    rng = np.random.default_rng(77 + padded_height)
    n, h = d["trace_len"], d["num_trace_randomizers"]
    assert n >= 512
    main, mrand = rand_bfes(rng, (379, n)), rand_bfes(rng, (379, h))
    qrand = rand_bfes(rng, (d["num_quotient_randomizer_coefficients"], 3))
    aux_t, aux_r = rand_bfes(rng, (91, n, 3)), rand_bfes(rng, (91, h, 3))
    claim = S.Claim([5, 4, 3, 2, 1], [], [])
    want, _ = S.prove(st, claim, main, mrand, lambda ch: (aux_t, aux_r), qrand, padded_height=padded_height)
    got = backend.prove((claim.program_digest, claim.input, claim.output), main, mrand, lambda ch: (aux_t, aux_r), qrand,
                        security_level=160, log2_expansion=2, padded_height=padded_height,
                        ldt_choice=tvm_b200.LDT_STIR if ldt == "stir" else tvm_b200.LDT_FRI)
    assert [int(v) for v in got] == want
    assert S.verify(st, claim, want, check_air=False)


def test_prove_reports_errors_instead_of_falling_back(backend):
    import tvm_b200
    st, d, claim, main, mrand, aux_provider, qrand = synthetic_instance(4, 2, 16, 9)
    args = ((claim.program_digest, claim.input, claim.output), main, mrand)
    with pytest.raises(RuntimeError, match="boom"):          # an exception inside the aux callback surfaces to the caller
        backend.prove(*args, lambda ch: (_ for _ in ()).throw(RuntimeError("boom")), qrand, security_level=4, log2_expansion=2,
                      padded_height=16)
    # unsupported expansion factor 64: explicit error, no fallback
    d1 = S.Stark(4, 6).derive(16)
    n1, h1 = d1["trace_len"], d1["num_trace_randomizers"]
    rng = np.random.default_rng(3)
    with pytest.raises(tvm_b200.TvmError) as e:
        backend.prove((claim.program_digest, claim.input, claim.output), rand_bfes(rng, (379, n1)), rand_bfes(rng, (379, h1)),
                      lambda ch: (rand_bfes(rng, (91, n1, 3)), rand_bfes(rng, (91, h1, 3))),
                      rand_bfes(rng, (d1["num_quotient_randomizer_coefficients"], 3)), security_level=4, log2_expansion=6, padded_height=16)
    assert e.value.code == -8


@pytest.mark.gpu
@pytest.mark.parametrize("ldt,security,padded_height", [("fri", 4, 16), ("stir", 6, 64)])
def test_host_kept_transcript_yields_the_same_proof(backend, ldt, security, padded_height):
    """tvm_prove_transcript (the host keeps ProofStream): the library only produces items and consumes challenges through
    callbacks; a host-side stream built on the oracle's sponge ends up holding exactly the proof tvm_prove returns."""
    import tvm_b200
    from oracle import codec

    class HostStream(codec.ProofStream):           # the Rust host's ProofStream, played by the oracle's
        def __init__(self):
            super().__init__()
            self.raw = []

        def enqueue_raw(self, variant, payload):
            e = [variant] + ([len(payload)] if variant > 4 else []) + list(payload)      # dynamically sized payloads carry their length
            if variant <= 6:                                                             # proof_item.rs:96-147: in the Fiat-Shamir heuristic
                self.alter_fiat_shamir_state_with(e)
            self.raw.append(e)

        def encode(self):
            body = [len(self.raw)]
            for e in self.raw:
                body += [len(e)] + e
            return [len(body)] + body

    rng = np.random.default_rng(21)
    st = S.Stark(security, 2, ldt)
    d = st.derive(padded_height)
    n, h = d["trace_len"], d["num_trace_randomizers"]
    main, mrand = rand_bfes(rng, (379, n)), rand_bfes(rng, (379, h))
    aux, arand = rand_bfes(rng, (91, n, 3)), rand_bfes(rng, (91, h, 3))
    qrand = rand_bfes(rng, (d["num_quotient_randomizer_coefficients"], 3))
    claim = ([1, 2, 3, 4, 5], [6], [7])
    choice = tvm_b200.LDT_STIR if ldt == "stir" else tvm_b200.LDT_FRI
    want = backend.prove(claim, main, mrand, lambda ch: (aux, arand), qrand, security_level=security, log2_expansion=2,
                         padded_height=padded_height, ldt_choice=choice)
    hs = HostStream()
    backend.prove_transcript(claim, main, mrand, lambda ch: (aux, arand), qrand, hs, security_level=security, log2_expansion=2,
                             padded_height=padded_height, ldt_choice=choice)
    assert hs.encode() == [int(v) for v in want]

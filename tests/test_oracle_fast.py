"""oracle/fast.py (every array stage in C/OpenMP: the complete CPU prove that bench.py --impl reference times) produces the
same proof words as oracle/stark.py's Python-int restatement, FRI and STIR."""
import numpy as np
import pytest

from oracle import fast, stark as S
from conftest import rand_bfes


@pytest.mark.parametrize("security,ldt,padded_height", [(4, "fri", 16), (6, "stir", 64), (8, "stir", 256)])
def test_fast_prove_equals_python_prove(security, ldt, padded_height):
    rng = np.random.default_rng(11)
    st = S.Stark(security, 2, ldt)
    d = st.derive(padded_height)
    n, h = d["trace_len"], d["num_trace_randomizers"]
    main, mrand = rand_bfes(rng, (379, n)), rand_bfes(rng, (379, h))
    aux, arand = rand_bfes(rng, (91, n, 3)), rand_bfes(rng, (91, h, 3))
    qrand = rand_bfes(rng, (d["num_quotient_randomizer_coefficients"], 3))
    claim = S.Claim([1, 2, 3, 4, 5], [6], [7])
    want, _ = S.prove(st, claim, main, mrand, lambda ch: (aux, arand), qrand, padded_height=padded_height)
    timings = {}
    got = fast.prove(st, claim, main, mrand, lambda ch: (aux, arand), qrand, padded_height=padded_height, timings=timings)
    assert got == want
    assert S.verify(st, claim, got, check_air=False)
    assert "low-degree test" in timings and all(v >= 0 for v in timings.values())

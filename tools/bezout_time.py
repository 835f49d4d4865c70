#!/usr/bin/env python3
"""Times tvm_bezout_coefficients (ram.rs:162-214 on the device) for m unique RAM pointers; checks a rp + b rp' = 1 at one point.
    python tools/bezout_time.py 16 18 20"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "triton-vm_b200", "py")]
import tvm_b200   # noqa: E402

P = tvm_b200.P
b = tvm_b200.Backend(0)
for lg in [int(v) for v in sys.argv[1:]] or [16]:
    m = 1 << lg
    rng = np.random.default_rng(lg)
    roots = np.unique(np.concatenate([np.arange(1 << 32, (1 << 32) + m // 2, dtype=np.uint64), rng.integers(0, P, size=m // 2, dtype=np.uint64)]))
    b.bezout_coefficients(roots[:1000])
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); a_, b_ = b.bezout_coefficients(roots); ts.append((time.perf_counter() - t0) * 1e3)
    z, ok = 3, None
    if lg <= 18:                                   # the identity at one point (python ints: O(m))
        rp, fd = 1, 0
        for r in (int(v) for v in roots):
            fd = (fd * (z - r) + rp) % P
            rp = rp * (z - r) % P
        ev = lambda poly: int(sum(int(c) * pow(z, k, P) for k, c in enumerate(poly)) % P)    # noqa: E731
        ok = (ev(a_) * rp + ev(b_) * fd) % P == 1
    print("m = 2^%d (%d roots): %s ms (incl. host copies), identity %s" % (lg, roots.size, [round(t, 1) for t in ts], ok), flush=True)

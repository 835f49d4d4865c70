#!/usr/bin/env python3
"""Times MasterMainTable::new + pad on the device (tvm_main_table_from_aet) on a SYNTHETIC AlgebraicExecutionTrace of 2^k
processor rows: uniform field elements in the traces, clk = row index, `ram_frac` of the rows touch RAM with half of the
pointers unique — the stages do not look at the semantics of the rows, only at their sizes and sort keys.  For the launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file fill.csv python tools/fill_time.py 20
    python tools/fill_time.py 16 18 20"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "triton-vm_b200", "py")]
import tvm_b200   # noqa: E402

P = tvm_b200.P


def synthetic_aet(k, rng, ram_frac=0.5, hash_frac=0.25):
    n = 1 << k
    plen = n - 5
    proc = rng.integers(0, P, size=(plen, 39), dtype=np.uint64)
    proc[:, 0] = np.arange(plen)
    proc[:, 12] = np.minimum(np.arange(plen) // 1000, rng.integers(0, 40, size=plen)).astype(np.uint64)   # JSP: a few dozen groups
    nos = n // 2
    ops = rng.integers(0, P, size=(nos, 4), dtype=np.uint64)
    ops[:, 0] = np.sort(rng.integers(0, plen, size=nos)); ops[:, 1] = rng.integers(0, 2, size=nos); ops[:, 2] = 16 + rng.integers(0, 5000, size=nos)
    nram = int(n * ram_frac)
    ram = np.zeros((nram, 7), dtype=np.uint64)
    ram[:, 0] = np.sort(rng.integers(0, plen, size=nram)); ram[:, 1] = rng.integers(0, 2, size=nram)
    uniq = rng.integers(0, P, size=nram // 2, dtype=np.uint64)
    ram[:, 2] = uniq[rng.integers(0, uniq.size, size=nram)]; ram[:, 3] = rng.integers(0, P, size=nram, dtype=np.uint64)
    nh = int(n * hash_frac)
    hrows = lambda r: rng.integers(0, P, size=(r, 67), dtype=np.uint64)       # noqa: E731
    nu = n // 48
    u32 = np.zeros((nu, 4), dtype=np.uint64)
    u32[:, 0] = rng.choice([4, 6, 14, 12, 30, 28], size=nu); u32[:, 1] = rng.integers(1, 1 << 32, size=nu); u32[:, 2] = rng.integers(0, 1 << 32, size=nu)
    u32[:, 3] = 1
    nc = min(60000, n // 2)
    casc = np.stack([np.arange(nc, dtype=np.uint64) % 65536, rng.integers(1, 1000, size=nc).astype(np.uint64)], axis=1)
    return dict(program=rng.integers(0, 100, size=1000, dtype=np.uint64), instruction_multiplicities=rng.integers(0, 1000, size=1000).astype(np.uint32),
                processor_trace=proc, op_stack_underflow_trace=ops, ram_trace=ram, program_hash_trace=hrows(606), sponge_trace=hrows(nh // 2),
                hash_trace=hrows(nh // 2), u32_entries=u32, cascade_table_lookup_multiplicities=casc,
                lookup_table_lookup_multiplicities=rng.integers(0, 1000, size=256).astype(np.uint64))


b = tvm_b200.Backend(0)
for k in [int(v) for v in sys.argv[1:]] or [16]:
    aet = synthetic_aet(k, np.random.default_rng(k))
    b.main_table_from_aet(aet, 1 << k)
    l0 = b.launches
    t0 = time.perf_counter()
    table, lengths = b.main_table_from_aet(aet, 1 << k)
    ms = (time.perf_counter() - t0) * 1e3
    print("2^%d rows: %.1f ms wall (incl. the AET upload from pageable memory and 379 columns back to the host), %d launches, table lengths %s"
          % (k, ms, b.launches - l0, lengths), flush=True)

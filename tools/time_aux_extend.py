#!/usr/bin/env python3
"""Wall-clock of the device-side table stages through the C ABI (synchronous calls): tvm_aux_extend and
tvm_fill_derived_main_columns on device-resident tensors, after warm-up, at the bench's padded height.
The first measurement the round-1 build could not take (its GPU minutes ended with the parity run).

    python tools/time_aux_extend.py [--log2-height 20] [--reps 5]           # TVM_AUX_TOPS_PARALLEL=1 for the variant
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "triton-vm_b200", "py")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2-height", type=int, default=20)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    import torch
    import tvm_b200
    n = 1 << a.log2_height
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1)
    P = tvm_b200.P
    main_t = (torch.randint(0, 2 ** 62, (379, n), generator=g, device=dev, dtype=torch.int64) % (P - 2 ** 63)).contiguous()   # < p
    ch = torch.randint(0, 2 ** 62, (63, 3), generator=g, dtype=torch.int64, device=dev).cpu().numpy().astype("uint64")
    out = torch.empty((91, n, 3), dtype=torch.int64, device=dev)
    b = tvm_b200.Backend(0)
    for name, fn in (("tvm_aux_extend", lambda: b.aux_extend(main_t, ch, None, out)),
                     ("tvm_fill_derived_main_columns", lambda: b.fill_derived_main_columns(main_t))):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(a.reps):
            t = time.perf_counter()
            fn()                                            # the C entry points synchronise their stream before returning
            ts.append(1e3 * (time.perf_counter() - t))
        print(f"{name}: n = 2^{a.log2_height}: min {min(ts):.1f} ms, median {sorted(ts)[len(ts) // 2]:.1f} ms "
              f"(includes the device-to-device staging copies and Montgomery conversions)")


if __name__ == "__main__":
    main()

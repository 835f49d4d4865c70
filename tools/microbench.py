"""Kernel microbenchmarks on one B200 (CUDA events on the library's stream)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "triton-vm_b200", "py"))
import numpy as np, torch
import tvm_b200

def timeit(fn, warm=2, it=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / it

def main():
    dev = torch.device("cuda:0")
    b = tvm_b200.Backend(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)   # events below are recorded on this (non-default) stream
    b.set_stream(stream.cuda_stream)
    res = {}
    g = torch.Generator(device=dev); g.manual_seed(1)
    def rnd(*shape):
        return torch.randint(0, 2**62, shape, dtype=torch.int64, device=dev, generator=g)
    # NTT
    for log2n, ncols in [(16, 256), (20, 64), (23, 8)]:
        n = 1 << log2n
        x = rnd(ncols, n); o = torch.empty_like(x); t = torch.empty_like(x)
        ms = timeit(lambda: b.ntt_dev(x, o, t, log2n, ncols))
        res[f"ntt_2^{log2n}x{ncols}"] = dict(ms=ms, gelem_s=ncols * n / ms / 1e6, gbs_alg=16 * ncols * n / ms / 1e6)
    # LDE
    for log2t, ncols in [(16, 64), (20, 16)]:
        n = 1 << log2t; rn = n * 8; h = 228
        tr = rnd(ncols, n); rd = rnd(ncols, h)
        coef = torch.empty((ncols, 2 * n), dtype=torch.int64, device=dev)
        out = torch.empty((ncols, rn), dtype=torch.int64, device=dev); tmp = torch.empty_like(out)
        ms = timeit(lambda: b.lde_dev(tr, rd, h, log2t, 3, 7, ncols, coef, out, tmp))
        res[f"lde_2^{log2t}x{ncols}"] = dict(ms=ms, gbs_alg=72 * ncols * n / ms / 1e6, out_gelem_s=ncols * rn / ms / 1e6)
    # hash rows
    for ncols, log2r in [(379, 20), (273, 20), (15, 22)]:
        nrows = 1 << log2r
        tab = rnd(ncols, nrows); dg = torch.empty((nrows, 5), dtype=torch.int64, device=dev)
        ms = timeit(lambda: b.hash_rows_dev(tab, nrows, nrows, ncols, 3, dg), it=3)
        perms = nrows * (ncols // 10 + 1)
        res[f"hash_rows_{ncols}x2^{log2r}"] = dict(ms=ms, mperm_s=perms / ms / 1e3, gbs_alg=(8 * ncols + 40) * nrows / ms / 1e6)
    # AIR quotient
    for log_n in (17,):
        log_r = 3; rn = 1 << (log_n + log_r)
        main = rnd(379, rn) % (2**61); aux = rnd(270, rn) % (2**61)
        out = torch.zeros((3, rn), dtype=torch.int64, device=dev)
        ch = np.arange(1, 190, dtype=np.uint64); w = np.arange(1, 1813, dtype=np.uint64)
        ms = timeit(lambda: b.air_quotient_dev(main, rn, aux, rn, ch, w, log_n, log_r, 7, out, rn), warm=1, it=3)
        res[f"air_quotient_2^{log_n+log_r}rows"] = dict(ms=ms, mrows_s=rn / ms / 1e3, gbs_alg=(5216 + 24) * rn / ms / 1e6)
    # merkle
    nl = 1 << 22
    nodes = rnd(2 * nl, 5)
    ms = timeit(lambda: b.merkle_dev(nodes, nl), it=3)
    res["merkle_2^22"] = dict(ms=ms, mperm_s=(nl - 1) / ms / 1e3)
    for k, v in res.items():
        print(k, json.dumps({a: round(c, 3) for a, c in v.items()}))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "microbench.json"), "w"), indent=1)

if __name__ == "__main__":
    main()

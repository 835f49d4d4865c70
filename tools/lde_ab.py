"""LDE timing only (tools/microbench.py's LDE section): 16 columns 2^20 -> 2^23, CUDA events on the library stream"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "triton-vm_b200", "py")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import tvm_b200
from microbench import timeit
dev = torch.device("cuda:0")
b = tvm_b200.Backend(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); b.set_stream(stream.cuda_stream)
g = torch.Generator(device=dev); g.manual_seed(1)
rnd = lambda *shape: torch.randint(0, 2**62, shape, dtype=torch.int64, device=dev, generator=g)
for log2t, ncols in [(20, 16), (18, 16), (16, 64)]:
    n = 1 << log2t; rn = n * 8; h = 228
    tr = rnd(ncols, n); rd = rnd(ncols, h)
    coef = torch.empty((ncols, 2 * n), dtype=torch.int64, device=dev)
    out = torch.empty((ncols, rn), dtype=torch.int64, device=dev); tmp = torch.empty_like(out)
    ms = timeit(lambda: b.lde_dev(tr, rd, h, log2t, 3, 7, ncols, coef, out, tmp), warm=3, it=10)
    print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("TVM_NTT")}, "lde": f"2^{log2t}x{ncols}", "ms": round(ms, 3),
                      "ms_652_cols": round(ms * 652 / ncols, 1), "checksum": int(out.sum().item())}))

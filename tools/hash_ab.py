"""Row-hashing timing only (tools/microbench.py's section): main / aux / quotient table shapes at 2^20 rows"""
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
for ncols, log2r in [(379, 20), (273, 20), (15, 22)]:
    nrows = 1 << log2r
    tab = rnd(ncols, nrows); dg = torch.empty((nrows, 5), dtype=torch.int64, device=dev)
    ms = timeit(lambda: b.hash_rows_dev(tab, nrows, nrows, ncols, 3, dg), warm=2, it=5)
    perms = nrows * (ncols // 10 + 1)
    print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("TVM_TIP5")}, "hash_rows": f"{ncols}x2^{log2r}", "ms": round(ms, 3),
                      "gperm_s": round(perms / ms / 1e6, 3), "checksum": int(dg.sum().item())}))

"""A/B timing of the generated AIR quotient kernels: python tools/air_ab.py  (TVM_B200_LIB selects the library build)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "triton-vm_b200", "py"))
import numpy as np, torch
import tvm_b200
from microbench import timeit

def main():
    dev = torch.device("cuda:0")
    b = tvm_b200.Backend(0)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); b.set_stream(stream.cuda_stream)
    g = torch.Generator(device=dev); g.manual_seed(1)
    rnd = lambda *shape: torch.randint(0, 2**61, shape, dtype=torch.int64, device=dev, generator=g)
    log_n, log_r = 18, 3
    rn = 1 << (log_n + log_r)
    main_t, aux_t = rnd(379, rn), rnd(270, rn)
    out = torch.zeros((3, rn), dtype=torch.int64, device=dev)
    ch = np.arange(1, 190, dtype=np.uint64); w = np.arange(1, 1813, dtype=np.uint64)
    ms = timeit(lambda: b.air_quotient_dev(main_t, rn, aux_t, rn, ch, w, log_n, log_r, 7, out, rn), warm=1, it=3)
    print(json.dumps({"lib": os.environ.get("TVM_B200_LIB", "default"), "rows": rn, "ms": round(ms, 3), "ms_at_2^23": round(ms * (1 << 23) / rn, 1),
                      "checksum": int(out.sum().item())}))

if __name__ == "__main__":
    main()

"""tvm_stir_prove alone at the size of a 2^20 prove (codeword of 2^23 X-field elements): wall clock per call"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "triton-vm_b200", "py"))
import numpy as np, torch
import tvm_b200
b = tvm_b200.Backend(0)
hdb = int(sys.argv[1]) if len(sys.argv) > 1 else 21
g = torch.Generator(device="cuda"); g.manual_seed(1)
cw = torch.randint(0, 2**62, ((1 << (hdb + 2)), 3), dtype=torch.int64, device="cuda", generator=g)
for _ in range(2):
    b.stir_prove(160, 2, hdb, cw)
torch.cuda.synchronize()
l0 = b.launches
t0 = time.perf_counter()
for _ in range(3):
    proof, idx = b.stir_prove(160, 2, hdb, cw)
torch.cuda.synchronize()
print(json.dumps({"log2_high_degree_bound": hdb, "ms_per_call": (time.perf_counter() - t0) / 3 * 1e3, "launches_per_call": (b.launches - l0) / 3, "proof_words": int(proof.size)}))

"""Small invocations of every round-2 kernel family for compute-sanitizer (memcheck / racecheck): tile NTT at R = 8, 16, 32 (LDE of a few
columns), row hashing through the TMA-staged and the LDG kernel, Merkle levels, a complete small prove (FRI + STIR), the device
table stages.  Results are compared with each other / the verifier so a silent corruption shows as well."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "triton-vm_b200", "py"))
import numpy as np
import tvm_b200
P = tvm_b200.P
b = tvm_b200.Backend(0)
rng = np.random.default_rng(3)
r = lambda *s: rng.integers(0, P, size=s, dtype=np.uint64)
for log2t, ncols in ((12, 3), (14, 2), (16, 2), (18, 1), (20, 1)):
    n = 1 << log2t
    out = b.lde(r(ncols, n), r(ncols, 40), 3, 7)
    print("lde 2^%d ok" % log2t, out.shape, int(out[0, 5]))
tab = r(25, 4096)
d1 = b.hash_rows(tab)
print("hash_rows (TMA path) ok", d1.shape)
tab2 = r(25, 4096 + 40)            # not a multiple of 64: LDG kernel
print("hash_rows (LDG path) ok", b.hash_rows(tab2).shape)
print("merkle ok", b.merkle(d1)[:2])
for sec, ldt, ph in ((4, tvm_b200.LDT_FRI, 16), (6, tvm_b200.LDT_STIR, 64)):
    dom = tvm_b200.derive_domains(sec, 2, ph, ldt)
    n, h = dom["trace_len"], dom["num_trace_randomizers"]
    claim = ([1, 2, 3, 4, 5], [6], [7])
    aux, arand = r(91, n, 3), r(91, h, 3)
    proof = b.prove(claim, r(379, n), r(379, h), lambda ch: (aux, arand), r(dom["num_quotient_randomizer_coefficients"], 3), sec, 2, ph, ldt)
    ok, why = tvm_b200.verify(claim, proof, sec, 2, ldt_choice=ldt, skip_air_check=True)
    assert ok, why
    print("prove + verify ok", sec, ph, proof.size)
main = r(379, 4096); col90 = r(4096, 3)
aux_t = b.aux_extend(main, r(63, 3), col90)
b.fill_derived_main_columns(main)
print("aux_extend / derived columns ok", aux_t.shape)
cw = r(1 << 12, 3)
proof, idx = b.stir_prove(12, 2, 10, cw)
print("stir_prove ok", proof.size, len(idx))
b.close()

"""GPU tests of kernels written AFTER the round's last GPU run (tvm_fill_derived_main_columns; the CTA-wide variant of the
chunk-total scan).  Their rules are checked on the host (tests/test_aux_extend.py), the kernels themselves have not run on
a GPU yet — the file sorts last so that, under `pytest -x`, a surprise here cannot hide the established GPU suites."""
import numpy as np
import pytest

import tvm_b200
from oracle import stark as S

from conftest import rand_bfes
from oracle import corc
from test_fibonacci_program import FIBONACCI, tables
from test_native_verifier import _build_example, _write_prove_tables_dir

pytestmark = pytest.mark.gpu


def test_device_fills_derived_main_columns(backend):
    main = tables(FIBONACCI, [7])[4]                         # a real table: the substitutions of live constraints
    got = main.copy()
    got[149:] = 7
    backend.fill_derived_main_columns(got)
    assert np.array_equal(got, main)
    rnd = rand_bfes(np.random.default_rng(3), (379, 1024))    # and arbitrary field elements in the 149 table columns
    want = corc.fill_derived_main(rnd)
    backend.fill_derived_main_columns(rnd)
    assert np.array_equal(rnd, want)


def test_parallel_scan_of_chunk_totals_variant():
    # TVM_AUX_TOPS_PARALLEL is read once per process: run both variants in fresh interpreters at n = 2^17 (512 chunks, two
    # per thread of the CTA-wide scan) and compare digests of the complete auxiliary table; the default (sequential) variant
    # is the one the other tests pin to the CPU rules
    import hashlib, subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = """
import sys, hashlib, numpy as np
sys.path[:0] = [%r, %r, %r]
import tvm_b200
from conftest import rand_bfes
rng = np.random.default_rng(5)
T = rand_bfes(rng, (379, 1 << 17)); ch = rand_bfes(rng, (63, 3))
print(hashlib.sha256(tvm_b200.Backend(0).aux_extend(T, ch).tobytes()).hexdigest())
""" % (root, os.path.join(root, "tests"), os.path.join(root, "triton-vm_b200", "py"))
    digests = [subprocess.run([sys.executable, "-c", code], check=True, capture_output=True, text=True, timeout=600,
                              env=dict(os.environ, TVM_AUX_TOPS_PARALLEL=flag)).stdout.strip().splitlines()[-1] for flag in ("0", "1")]
    assert len(digests[0]) == 64 and digests[0] == digests[1]


def test_plain_c_prover_client_proves_halt_on_the_gpu(tmp_path):
    import subprocess
    exe = _build_example(tmp_path, "prove_tables")
    digest = _write_prove_tables_dir(tmp_path)
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0 and "verified" in r.stdout, r.stderr
    proof = np.fromfile(str(tmp_path / "proof.u64"), dtype="<u8")
    assert tvm_b200.verify((digest, [], []), proof, 8, 2, ldt_choice=tvm_b200.LDT_FRI) == (True, "")
    assert S.verify(S.Stark(8, 2, "fri"), S.Claim(digest, [], []), [int(v) for v in proof], check_air=True)

"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/tvm_b200.h declares; host-only entry points compute; GPU entry points fail loudly
(no CPU fallback)."""
import ctypes
import os
import re

import pytest

import tvm_b200
from oracle import tip5 as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    return tvm_b200.build()


def test_library_exports_every_declared_symbol(built):
    header = open(os.path.join(ROOT, "include", "tvm_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(tvm_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 15
    l = ctypes.CDLL(built)
    for name in sorted(declared):
        assert hasattr(l, name), f"{name} declared in include/tvm_b200.h but not exported"
    # and the Python mirror binds all of them
    assert declared == set(tvm_b200._SIGNATURES), declared ^ set(tvm_b200._SIGNATURES)


def test_host_hash_varlen_matches_oracle(built):
    import random
    rng = random.Random(3)
    for n in [0, 1, 9, 10, 11, 64]:
        w = [rng.randrange(T.P) for _ in range(n)]
        assert tvm_b200.hash_varlen(w) == T.hash_varlen(w)


def test_error_strings(built):
    l = tvm_b200.lib()
    assert l.tvm_strerror(0) == b"ok"
    assert b"OutOfMemory" in l.tvm_strerror(-3)
    assert b"ZeroKnowledgeViolation" in l.tvm_strerror(-4)


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(tvm_b200.TvmError):
        tvm_b200.Backend(0)

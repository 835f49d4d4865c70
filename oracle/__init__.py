"""CPU oracle for the Triton VM `Stark::prove()` hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under `oracle/` is product code: only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import, link or execute it, and there only as the
checker / reported CPU baseline (plus the fixture and workload generators under `tests/golden/` and
`tools/make_workload.py`, which are developer tools).  The product path (`triton-vm_b200/`) never
routes through it and fails loudly when the CUDA library is missing.

The oracle restates the reference algorithm (TritonVM/triton-vm @ 8cd9a0eb) and
the published algorithms of its un-vendored dependency `twenty-first = "2.0.0"`
(reference `Cargo.toml:104`; field, NTT, Tip5, Merkle tree, BFieldCodec).
Parity: PINNED.  The oracle reproduces the reference's Tip5 and Montgomery known answers, the AIR fingerprint and — end to
end — both whole-proof known-answer digests (proof.rs:200-226, stark.rs:2433-2460; tests/test_golden.py).  Module map and
pin status: `oracle/README.md`.
"""

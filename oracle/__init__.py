"""CPU oracle for the Triton VM `Stark::prove()` hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under `oracle/` is product code: only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import, link or execute it, and there only as the
checker / reported CPU baseline.  The product path (`triton-vm_b200/`) never
routes through it and fails loudly when the CUDA library is missing.

The oracle restates the reference algorithm (TritonVM/triton-vm @ 8cd9a0eb) and
the published algorithms of its un-vendored dependency `twenty-first = "2.0.0"`
(reference `Cargo.toml:104`; field, NTT, Tip5, Merkle tree, BFieldCodec).
Parity pins: see `oracle/PINNING.md`.
"""

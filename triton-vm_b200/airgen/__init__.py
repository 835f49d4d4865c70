"""Build-time AIR pipeline for libtvm_b200: a Python restatement of the reference's symbolic
AIR (triton-air), constraint circuit builder and degree lowering
(triton-constraint-circuit, triton-constraint-builder), ending in generated sm_100a CUDA source
for the quotient kernel and in a plain-data description of the circuit that the CPU oracle
evaluates node by node.

The reference generates its evaluator at build time too (triton-vm/build.rs:11-25) and does not
check it in, so the circuit has to be re-derived; everything here cites the file:line it follows.
"""

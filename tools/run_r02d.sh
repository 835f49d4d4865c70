#!/bin/bash
cd "$(dirname "$0")/.."
for v in chunks100 chunks160 groups220; do
  TVM_B200_LIB=$PWD/triton-vm_b200/lib/variants/libtvm_b200_$v.so python tools/air_ab.py 2>&1 | tail -1
done > gpurun_out/r02d_air_ab.log
cat gpurun_out/r02d_air_ab.log

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in default split400 split250; do
  if [ $v = default ]; then python tools/air_ab.py 2>&1 | tail -1; else TVM_B200_LIB=$PWD/triton-vm_b200/lib/variants/libtvm_b200_$v.so python tools/air_ab.py 2>&1 | tail -1; fi
done > gpurun_out/r02o_air_ab.log
python tools/stir_profile.py 21 > gpurun_out/r02o_stir.log 2>&1
python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "merkle" > gpurun_out/r02o_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02o_tests.log
cat gpurun_out/r02o_air_ab.log; tail -1 gpurun_out/r02o_stir.log; tail -2 gpurun_out/r02o_tests.log

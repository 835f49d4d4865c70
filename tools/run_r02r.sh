#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prove.py tests/test_gpu_sharded_prove.py tests/test_stir_properties.py -m gpu -q -x > gpurun_out/r02r_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02r_tests.log
tail -6 gpurun_out/r02r_tests.log

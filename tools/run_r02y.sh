#!/bin/bash
# r02y: device-table modes (tvm_prove_tables / tvm_prove_aet) sharded over several ranks; regression of the sharded suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_sharded_prove.py tests/test_z_gpu_device_tables.py tests/test_zz_gpu_main_fill.py -x -q -m gpu > gpurun_out/r02y_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02y_tests.log
tail -25 gpurun_out/r02y_tests.log

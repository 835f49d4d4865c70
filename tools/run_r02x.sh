#!/bin/bash
# r02x: A/B of the single-128-bit-product MAC in the AIR accumulation (air.cuh) and the dot products: parity tests + bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_prove.py -x -q -m gpu > gpurun_out/r02x_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02x_tests.log
tail -3 gpurun_out/r02x_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.err
tail -1 gpurun_out/r02x_bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['stages_ms'], d['proof_check'])"

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_golden.py -m gpu -q -x > gpurun_out/r02p_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02p_tests.log
timeout 120 python tools/hash_ab.py > gpurun_out/r02p_hash_tma.log 2>&1; echo "rc=$?" >> gpurun_out/r02p_hash_tma.log
TVM_TIP5_NO_TMA=1 timeout 120 python tools/hash_ab.py > gpurun_out/r02p_hash_ldg.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err
tail -4 gpurun_out/r02p_tests.log; grep hash_rows gpurun_out/r02p_hash_tma.log gpurun_out/r02p_hash_ldg.log | cut -c1-200
tail -1 gpurun_out/r02p_bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['stages_ms'], d['proof_check']['accepted'])"

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_golden.py -m gpu -q -x > gpurun_out/r02q_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02q_tests.log
timeout 200 python tools/lde_ab.py 2>&1 | grep lde > gpurun_out/r02q_lde_tma.log
TVM_NTT_NO_TMA=1 timeout 200 python tools/lde_ab.py 2>&1 | grep lde > gpurun_out/r02q_lde_ldg.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err
tail -4 gpurun_out/r02q_tests.log; cat gpurun_out/r02q_lde_tma.log gpurun_out/r02q_lde_ldg.log | cut -c1-200
tail -1 gpurun_out/r02q_bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['stages_ms'], d['proof_check']['accepted'])"

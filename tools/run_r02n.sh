#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_sharded_prove.py tests/test_stir_properties.py tests/test_golden.py tests/test_gpu_prove.py -m gpu -q -x > gpurun_out/r02n_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02n_tests.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err
python tools/stir_profile.py 21 > gpurun_out/r02n_stir.log 2>&1
tail -4 gpurun_out/r02n_tests.log; tail -1 gpurun_out/r02n_stir.log
tail -1 gpurun_out/r02n_bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['stages_ms'], d['proof_check']['accepted'])"

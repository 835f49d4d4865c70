#!/bin/bash
# r02aa: degree split of the AIR quotient (constraints of degree <= 2 on every second coset + extension): parity and timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prove.py tests/test_gpu_kernels.py tests/test_vm_programs.py -x -q -m gpu -k "not spin_18 and not spin_20" > gpurun_out/r02aa_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02aa_tests.log
tail -6 gpurun_out/r02aa_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02aa_bench.json 2> gpurun_out/r02aa_bench.err
tail -1 gpurun_out/r02aa_bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['stages_ms'], d['proof_check'])"
TVM_AIR_NO_DEGREE_SPLIT=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02aa_bench_nosplit.json 2> gpurun_out/r02aa_bench_nosplit.err
tail -1 gpurun_out/r02aa_bench_nosplit.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['stages_ms']['quotient(AIR)'])"

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_aux_extend.py tests/test_z_gpu_device_tables.py tests/test_vm_programs.py -m gpu -q -x > gpurun_out/r02j_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02j_tests.log
python tools/make_workload.py spin_20 /tmp/spin20 > /dev/null 2>&1
python bench.py --workload-dir /tmp/spin20 --steps 5 --warmup 2 > gpurun_out/r02j_bench_spin20.json 2> gpurun_out/r02j_bench_spin20.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err
tail -3 gpurun_out/r02j_tests.log
for f in r02j_bench_spin20 r02j_bench; do tail -1 gpurun_out/$f.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['stages_ms'], d.get('proof_check',{}).get('accepted'))"; done

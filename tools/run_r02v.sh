#!/bin/bash
# r02v: first GPU run of the device table fill (tvm_main_table_from_aet / tvm_prove_aet / tvm_bezout_coefficients)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_gpu_main_fill.py "tests/test_gpu_prove.py::test_low_memory_mode_produces_the_same_proof" -x -q > gpurun_out/r02v_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02v_tests.log
tail -15 gpurun_out/r02v_tests.log
timeout 300 python tools/bezout_time.py 14 16 18 20 > gpurun_out/r02v_bezout.log 2>&1; cat gpurun_out/r02v_bezout.log
timeout 600 python tools/make_workload.py spin_20 /tmp/spin20 --aet > gpurun_out/r02v_make.log 2>&1; tail -2 gpurun_out/r02v_make.log
timeout 300 python bench.py --workload-dir /tmp/spin20 --from-aet --steps 5 --warmup 2 > gpurun_out/r02v_bench_spin20_aet.json 2> gpurun_out/r02v_bench_spin20_aet.err
timeout 300 python bench.py --workload-dir /tmp/spin20 --steps 5 --warmup 2 > gpurun_out/r02v_bench_spin20_tables.json 2> gpurun_out/r02v_bench_spin20_tables.err
for f in r02v_bench_spin20_aet r02v_bench_spin20_tables; do tail -1 gpurun_out/$f.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e']['h2d_bytes_per_step'], d['stages_ms'], d['proof_check']['accepted'])"; tail -3 gpurun_out/$f.err; done

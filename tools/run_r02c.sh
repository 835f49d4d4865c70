#!/bin/bash
# r02c: fused AIR groups + subtraction peephole: parity, bench; ncu --set full of the new kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_golden.py tests/test_gpu_prove.py -m gpu -q -x > gpurun_out/r02c_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02c_tests.log
python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r02c_bench.log 2>&1
B20="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:ntt_tile_kernel -s 24 -c 4 -f -o gpurun_out/r02c_ntt_tile $B20 > gpurun_out/r02c_ncu_a.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:tip5_hash_rows_mma -s 1 -c 1 -f -o gpurun_out/r02c_tip5_mma $B20 > gpurun_out/r02c_ncu_b.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:air_group_0[45]_tran -s 2 -c 2 -f -o gpurun_out/r02c_air_group $B20 > gpurun_out/r02c_ncu_c.log 2>&1
tail -3 gpurun_out/r02c_tests.log; tail -1 gpurun_out/r02c_bench.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['stages_ms'])"
ls -la gpurun_out | grep r02c

#!/bin/bash
# r02f: NTT tile A/B (looped rounds), DRAM traffic of the tile passes with the L2 left alone, ncu of the table-stage kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/microbench.py > gpurun_out/r02f_micro.log 2>&1
TVM_NTT_LOOPED_ROUNDS=1 python tools/microbench.py > gpurun_out/r02f_micro_looped.log 2>&1
grep -h "lde_" gpurun_out/r02f_micro.log gpurun_out/r02f_micro_looped.log
B20="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"
timeout 300 ncu --metrics $M --cache-control none --clock-control none -k regex:ntt_tile_kernel -s 40 -c 8 --csv --log-file gpurun_out/r02f_ntt_tile_traffic.csv $B20 > gpurun_out/r02f_ncu_a.log 2>&1
python tools/make_workload.py spin_18 /tmp/spin18 > gpurun_out/r02f_workload.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"aux_scan_kernel|aux_derived_kernel|main_derived_kernel|aux_scan_tops" -c 8 -f -o gpurun_out/r02f_aux_kernels python bench.py --workload-dir /tmp/spin18 --steps 1 --warmup 0 > gpurun_out/r02f_ncu_b.log 2>&1
grep -v "^==" gpurun_out/r02f_ntt_tile_traffic.csv | tail -50 | cut -c1-220

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in default minb2 minb3 budget60; do
  if [ $v = default ]; then python tools/air_ab.py 2>&1 | tail -1; else TVM_B200_LIB=$PWD/triton-vm_b200/lib/variants/libtvm_b200_$v.so python tools/air_ab.py 2>&1 | tail -1; fi
done > gpurun_out/r02i_air_ab.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r02i_launches_2p20.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02i_launches.log 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__inst_executed.sum"
python tools/make_workload.py spin_18 /tmp/spin18 > /dev/null 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:"aux_scan|aux_derived|main_derived" -c 12 --csv --log-file gpurun_out/r02i_aux_kernels_2p18.csv python bench.py --workload-dir /tmp/spin18 --steps 1 --warmup 0 > gpurun_out/r02i_aux.log 2>&1
cat gpurun_out/r02i_air_ab.log
tail -1 gpurun_out/r02i_bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['stages_ms'], d['roofline']['frac'])"
gzip -f gpurun_out/r02i_launches_2p20.csv; ls -la gpurun_out | grep r02i

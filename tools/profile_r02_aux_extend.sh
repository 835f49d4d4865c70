# First captures for round 2: the device-side extension (csrc/aux_extend.cu) has parity but no measurement yet.
set -x
T="python tools/time_aux_extend.py --log2-height 20 --reps 3"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum"
$T > gpurun_out/r02_aux_extend_times.txt 2>&1
TVM_AUX_TOPS_PARALLEL=1 $T >> gpurun_out/r02_aux_extend_times.txt 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:"aux_|main_derived" -c 40 --csv --log-file gpurun_out/r02_aux_extend_kernels.csv python tools/time_aux_extend.py --log2-height 20 --reps 1 > gpurun_out/r02_aux_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:aux_scan_kernel -s 2 -c 1 -f -o gpurun_out/r02_aux_scan_kernel python tools/time_aux_extend.py --log2-height 18 --reps 1 > gpurun_out/r02_aux_full.log 2>&1
ls -la gpurun_out | grep r02

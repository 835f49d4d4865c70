# Final captures of round 1.  Keep everything small: gpurun copies back at most 64 MiB and a `--set full` capture of
# every launch of a prove() takes > 40 min (memory save/restore per replay pass) - the all-kernels table therefore
# uses a short explicit metric list (2 passes).
set -x
B20="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
B18="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --log2-height 18"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread"
# 1) the LDE passes of a 16-column batch at 2^20 (DRAM traffic per launch for the roofline line)
timeout 300 ncu --set full --clock-control none -k regex:ntt_pass_a_kernel -s 40 -c 2 -f -o gpurun_out/r01d_ntt_pass_a_2p20 $B20 > gpurun_out/r01d_a.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:ntt_pass_b_kernel -s 40 -c 2 -f -o gpurun_out/r01d_ntt_pass_b_2p20 $B20 > gpurun_out/r01d_b.log 2>&1
# 2) launch list of the default bench command (2^20)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 7000 --csv --log-file gpurun_out/r01d_launches_2p20.csv $B20 > gpurun_out/r01d_launches_bench.log 2>&1
# 3) every kernel of one prove at 2^18, short metric list
timeout 500 ncu --metrics $M --clock-control none -c 2300 --csv --log-file gpurun_out/r01d_all_kernels_2p18.csv $B18 > gpurun_out/r01d_all.log 2>&1
ls -la gpurun_out | grep r01d

set -x
B20="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
B16="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --log2-height 16"
# 1) launch list of the default bench command (2^20)
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r01d_launches_2p20.csv $B20 > gpurun_out/r01d_launches_bench.log 2>&1
# 2) every kernel of one prove at 2^16, full metric set
ncu --set full --clock-control none -c 2300 -f -o gpurun_out/r01d_all_kernels_2p16 $B16 > gpurun_out/r01d_all.log 2>&1
# 3) the LDE passes of a 16-column batch at 2^20 (DRAM traffic per launch for the roofline line)
ncu --set full --clock-control none --import-source on -k regex:ntt_pass_a_kernel -s 40 -c 2 -f -o gpurun_out/r01d_ntt_pass_a_2p20 $B20 > gpurun_out/r01d_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ntt_pass_b_kernel -s 40 -c 2 -f -o gpurun_out/r01d_ntt_pass_b_2p20 $B20 > gpurun_out/r01d_b.log 2>&1
ls -la gpurun_out | grep r01d

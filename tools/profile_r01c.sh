set -x
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --log2-height 18"
ncu --set full --clock-control none --import-source on -k regex:tip5_hash_rows_quad2_kernel -c 1 -f -o gpurun_out/r01c_tip5_quad2 $B > gpurun_out/r01c_a.log 2>&1
TVM_TIP5_TWO_LANES=1 ncu --set full --clock-control none --import-source on -k regex:tip5_hash_rows_pair_kernel -c 1 -f -o gpurun_out/r01c_tip5_pair $B > gpurun_out/r01c_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ntt_pass_a_kernel -s 4 -c 1 -f -o gpurun_out/r01c_ntt_pass_a $B > gpurun_out/r01c_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ntt_pass_b_kernel -s 4 -c 1 -f -o gpurun_out/r01c_ntt_pass_b $B > gpurun_out/r01c_d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:air_chunk_060_tran -c 1 -f -o gpurun_out/r01c_air_chunk_060 $B > gpurun_out/r01c_e.log 2>&1
ls -la gpurun_out | tail -12

set -x
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r01b_launches_2p20.csv $B > gpurun_out/r01b_launches_bench.log 2>&1
for k in tip5_hash_rows_quad_kernel ntt_pass_a_kernel ntt_pass_b_kernel air_chunk_060_tran air_chunk_100_tran col_dot_kernel merkle_level_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/r01b_$k $B > gpurun_out/r01b_$k.log 2>&1
done
ls -la gpurun_out

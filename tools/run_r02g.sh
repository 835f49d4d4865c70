#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_stir_properties.py tests/test_golden.py -m gpu -q -x > gpurun_out/r02g_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02g_tests.log
for mb in 64 32 16; do TVM_NTT_TMP_MB=$mb python tools/lde_ab.py 2>&1 | grep lde; done > gpurun_out/r02g_lde_ab.log
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"
for mb in 32 16; do
TVM_NTT_TMP_MB=$mb timeout 300 ncu --metrics $M --cache-control none --clock-control none -k regex:ntt_tile_kernel -s 60 -c 8 --csv --log-file gpurun_out/r02g_ntt_tile_traffic_$mb.csv python tools/lde_ab.py > gpurun_out/r02g_ncu_$mb.log 2>&1
done
tail -4 gpurun_out/r02g_tests.log; cat gpurun_out/r02g_lde_ab.log
python - <<'PY'
import csv, io
for mb in (32, 16):
    rows = [l for l in open(f"gpurun_out/r02g_ntt_tile_traffic_{mb}.csv") if l.startswith('"')]
    agg = {}
    for r in csv.DictReader(io.StringIO("".join(rows))):
        agg.setdefault((r["ID"], r["Kernel Name"][:44], r["Grid Size"]), {})[r["Metric Name"]] = r["Metric Value"]
    print("TMP_MB", mb)
    for k, v in agg.items():
        print(k, {m.split("__")[-1][:24]: x for m, x in v.items()})
PY

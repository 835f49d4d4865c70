#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "lde or ntt" > gpurun_out/r02h_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02h_tests.log
for cfg in "1 64" "2 64" "2 32" "1 128"; do set -- $cfg; TVM_NTT_STREAMS=$1 TVM_NTT_TMP_MB=$2 python tools/lde_ab.py 2>&1 | grep lde; done > gpurun_out/r02h_lde_ab.log
# DRAM traffic of the tile passes with the caches left alone and NO kernel replay (two metrics = one pass)
for cfg in "1 64" "2 64" "2 32"; do set -- $cfg
TVM_NTT_STREAMS=$1 TVM_NTT_TMP_MB=$2 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none -k regex:ntt_tile_kernel -s 100 -c 64 --csv --log-file gpurun_out/r02h_traffic_s$1_mb$2.csv python tools/lde_ab.py > /dev/null 2>&1
done
python tools/make_workload.py spin_16 /tmp/spin16 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"aux_scan_kernel|aux_derived_kernel|main_derived_kernel" -c 6 -f -o gpurun_out/r02h_aux_kernels python bench.py --workload-dir /tmp/spin16 --steps 1 --warmup 0 > gpurun_out/r02h_ncu_b.log 2>&1
tail -3 gpurun_out/r02h_tests.log; cat gpurun_out/r02h_lde_ab.log
python - <<'PY'
import csv, io, glob
for f in sorted(glob.glob("gpurun_out/r02h_traffic_*.csv")):
    rows = [l for l in open(f) if l.startswith('"')]
    tot = {}
    n = 0
    for r in csv.DictReader(io.StringIO("".join(rows))):
        tot[r["Metric Name"]] = tot.get(r["Metric Name"], 0) + float(r["Metric Value"].replace(",", ""))
        n += 1
    print(f, "launches", n // 2, {k: round(v / 1e6, 1) for k, v in tot.items()}, "MB; per column-pair (8 cosets):", round(sum(tot.values()) / 1e6 / max(1, n // 2) * 2 * (2 if 'mb32' in f else 1), 1))
PY
ls -la gpurun_out/r02h_aux_kernels.ncu-rep

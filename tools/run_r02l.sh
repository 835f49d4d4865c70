#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/stir_profile.py 21 > gpurun_out/r02l_stir.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02l_stir_launches.csv python tools/stir_profile.py 21 > /dev/null 2>&1
python tools/summarize_ncu.py launches gpurun_out/r02l_stir_launches.csv gpurun_out/r02l_stir_launches.md > /dev/null 2>&1
cat gpurun_out/r02l_stir.log | tail -2; head -30 gpurun_out/r02l_stir_launches.md; rm -f gpurun_out/r02l_stir_launches.csv

#!/bin/bash
# r02w: complete GPU suite + smoke with the device table fill in the library; fill timing and launch list on a synthetic AET
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02w_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/r02w_gputests.log
tail -4 gpurun_out/r02w_gputests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02w_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r02w_smoke.log
tail -2 gpurun_out/r02w_smoke.log
timeout 300 python tools/fill_time.py 16 18 20 > gpurun_out/r02w_fill_time.log 2>&1; cat gpurun_out/r02w_fill_time.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02w_fill_launches.csv python tools/fill_time.py 20 > gpurun_out/r02w_ncu.log 2>&1
tail -2 gpurun_out/r02w_ncu.log; wc -l gpurun_out/r02w_fill_launches.csv

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/r02u_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/r02u_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/r02u_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/r02u_racecheck.log
grep -E "ERROR SUMMARY|rc=|ok|Error|RACECHECK SUMMARY|hazard" gpurun_out/r02u_memcheck.log | tail -25
grep -E "ERROR SUMMARY|rc=|RACECHECK SUMMARY|hazard|Race" gpurun_out/r02u_racecheck.log | sort | uniq -c | tail -15

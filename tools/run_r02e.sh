#!/bin/bash
# r02e: tvm_prove_tables parity (benchmark workloads), spin_20 bench from files, default bench, reference arm on the box's host cores
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_vm_programs.py -m gpu -q -x > gpurun_out/r02e_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02e_tests.log
python tools/make_workload.py spin_20 /tmp/spin20 > gpurun_out/r02e_workload.log 2>&1
python bench.py --workload-dir /tmp/spin20 --steps 5 --warmup 2 > gpurun_out/r02e_bench_spin20.json 2> gpurun_out/r02e_bench_spin20.err
python bench.py --steps 5 --warmup 3 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02e_bench_reference.json 2> gpurun_out/r02e_bench_reference.err
tail -3 gpurun_out/r02e_tests.log; cat gpurun_out/r02e_workload.log | tail -2
for f in r02e_bench_spin20 r02e_bench r02e_bench_reference; do python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"], 1), d.get("e2e", {}).get("value"), d.get("stages_ms"), d.get("proof_check"), d.get("cpu_baseline", {}).get("sample"), d.get("measured"))
except Exception as e:
    print("$f", "FAILED", e); print(open("gpurun_out/$f.err").read()[-1500:])
PY
done

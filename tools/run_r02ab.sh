#!/bin/bash
# r02ab: final single-GPU state of round 2 (with the degree split of the AIR quotient): complete GPU suite, smoke, default bench (with cpu_baseline), spin_20 from the AET
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02ab_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/r02ab_gputests.log
tail -4 gpurun_out/r02ab_gputests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02ab_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r02ab_smoke.log
tail -2 gpurun_out/r02ab_smoke.log
timeout 900 python bench.py > gpurun_out/r02ab_bench.json 2> gpurun_out/r02ab_bench.err
tail -1 gpurun_out/r02ab_bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['stages_ms'], d.get('roofline'), d.get('cpu_baseline',{}).get('value'), d['proof_check']['accepted'])"

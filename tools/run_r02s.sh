#!/bin/bash
# r02s: final single-GPU state: complete GPU suite, smoke, default bench with cpu_baseline, spin_20 workload bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02s_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/r02s_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02s_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r02s_smoke.log
python bench.py > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err
python tools/make_workload.py spin_20 /tmp/spin20 > /dev/null 2>&1
python bench.py --workload-dir /tmp/spin20 --steps 5 --warmup 2 > gpurun_out/r02s_bench_spin20.json 2> gpurun_out/r02s_bench_spin20.err
tail -4 gpurun_out/r02s_gputests.log; tail -2 gpurun_out/r02s_smoke.log
for f in r02s_bench r02s_bench_spin20; do tail -1 gpurun_out/$f.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['stages_ms'], d.get('roofline',{}).get('frac'), d.get('cpu_baseline',{}).get('value'), d['proof_check']['accepted'])"; done

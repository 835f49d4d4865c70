#!/bin/bash
# r02m: the complete GPU suite + smoke (as the driver runs them), then the verifier-shaped program at 2^20 through tvm_prove_tables
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/r02m_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/r02m_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02m_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r02m_smoke.log
python tools/make_workload.py verifier_11500 /tmp/ver20 > gpurun_out/r02m_workload.log 2>&1
python bench.py --workload-dir /tmp/ver20 --steps 5 --warmup 2 > gpurun_out/r02m_bench_verifier20.json 2> gpurun_out/r02m_bench_verifier20.err
tail -16 gpurun_out/r02m_gputests.log; tail -3 gpurun_out/r02m_smoke.log; tail -1 gpurun_out/r02m_workload.log
tail -1 gpurun_out/r02m_bench_verifier20.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['stages_ms'], d['proof_check'], d['config']['workload'][:60])" || tail -5 gpurun_out/r02m_bench_verifier20.err

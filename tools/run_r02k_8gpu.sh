#!/bin/bash
# r02k (8 GPUs of one box): one proof sharded over 8 B200 at 2^20, and BASELINE config 4: padded height 2^22 sharded over 8 GPUs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
timeout 300 $T bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02k_bench_8gpu_2p20.json 2> gpurun_out/r02k_bench_8gpu_2p20.err
timeout 500 $T bench.py --gpus 8 --steps 2 --warmup 1 --log2-height 22 > gpurun_out/r02k_bench_8gpu_2p22.json 2> gpurun_out/r02k_bench_8gpu_2p22.err
for f in r02k_bench_8gpu_2p20 r02k_bench_8gpu_2p22; do tail -1 gpurun_out/$f.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$f', d['value'], d['e2e']['value'], d['stages_ms'], d.get('proof_check',{}).get('accepted'), d['config']['lde_tables'])" || tail -5 gpurun_out/$f.err; done

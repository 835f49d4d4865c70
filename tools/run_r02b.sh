#!/bin/bash
# r02b: Tip5 with the MDS on IMMA vs the round-1 IMAD kernel; parity tests; bench stage split
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_golden.py tests/test_gpu_prove.py -m gpu -q -x > gpurun_out/r02b_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02b_tests.log
python tools/microbench.py > gpurun_out/r02b_micro_mma.log 2>&1
TVM_TIP5_IMAD_MDS=1 python tools/microbench.py > gpurun_out/r02b_micro_imad.log 2>&1
python bench.py --steps 3 --warmup 1 > gpurun_out/r02b_bench.log 2>&1
tail -5 gpurun_out/r02b_tests.log; grep -h "hash_rows\|merkle" gpurun_out/r02b_micro_mma.log gpurun_out/r02b_micro_imad.log; tail -2 gpurun_out/r02b_bench.log

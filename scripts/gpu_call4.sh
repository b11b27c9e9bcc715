#!/bin/bash
# GPU session: encoder/PPO parity tests, bench, ncu --set full of the five tensor-core conv kernels.
TAG=${1:-r01s}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest"; timeout 500 python -m pytest tests/test_policy_gpu.py tests/test_ppo_gpu.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -30 | tee $OUT/pytest_gpu.txt
echo "== bench"; timeout 240 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline 2>$OUT/bench_err.txt | tee $OUT/bench.json | cut -c1-300
timeout 400 ncu --nvtx --nvtx-include "timed/" --set full --clock-control none --import-source on \
    -k regex:"mma_kernel" -c 5 -o $OUT/prof_mma python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log | cut -c1-200
ls $OUT

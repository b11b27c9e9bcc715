#!/bin/bash
# ncu --set full of the conv2 mma.sync kernels (one launch each) inside the bench's timed region.
TAG=${1:-r01q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 200 python -m pytest tests/test_ppo_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -5 | tee $OUT/pytest_ppo.txt
timeout 400 ncu --nvtx --nvtx-include "timed/" --set full --clock-control none --import-source on \
    -k regex:"${2:-mma_kernel}" -c ${3:-3} -o $OUT/prof_mma python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
tail -3 $OUT/ncu_full.log
ls -la $OUT

#!/bin/bash
# ncu --set full (with source counters) of one launch of each hot kernel inside the bench's timed region.
# Usage (under gpurun, one GPU): bash scripts/gpu_ncu_full.sh <tag>
TAG=${1:-r02n}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 500 ncu --nvtx --nvtx-include "timed/" --set full --clock-control none --import-source on \
    -k regex:"scan_raycast|grid_update_sparse|conv1_fwd_mma|conv2_fwd_mma|tc_gemm|wgrad_staged|conv2_dgrad_mma|conv1_wgrad_mma" -c 10 \
    -o $OUT/prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ppo > $OUT/ncu_full.log 2>&1
tail -3 $OUT/ncu_full.log
ls -la $OUT

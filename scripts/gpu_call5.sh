#!/bin/bash
TAG=${1:-r01t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest"; timeout 600 python -m pytest tests/test_policy_gpu.py tests/test_ppo_gpu.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -30 | tee $OUT/pytest_gpu.txt
for mode in 14 30; do
  echo "== bench GNBV_CONV2_TC=$mode"
  GNBV_CONV2_TC=$mode timeout 240 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline 2>$OUT/bench_err_$mode.txt | tee $OUT/bench_conv2mode$mode.json | cut -c1-260
done
GNBV_CONV2_TC=30 timeout 400 ncu --nvtx --nvtx-include "timed/" --set full --clock-control none --import-source on \
    -k regex:"wgrad_staged|dgrad_mma" -c 2 -o $OUT/prof_mma python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log | cut -c1-200
ls $OUT

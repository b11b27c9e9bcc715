#!/bin/bash
# GPU session: full parity suite + conv1 kernel variants (CUDA-core vs mma.sync) + eval chamfer stress.
TAG=${1:-r01r}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -40 | tee $OUT/pytest_gpu.txt
for mode in 0 1 3; do
  echo "== bench GNBV_CONV1_MMA=$mode"
  GNBV_CONV1_MMA=$mode timeout 240 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline 2>$OUT/bench_err_$mode.txt | tee $OUT/bench_conv1mode$mode.json | cut -c1-300
done
echo "== chamfer stress"; timeout 420 python scripts/chamfer_stress.py --out $OUT/chamfer_stress.json 2>&1 | tail -2 | cut -c1-900
ls $OUT

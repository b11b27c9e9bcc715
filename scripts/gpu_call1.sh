#!/bin/bash
# GPU session A: full parity suite, conv2 kernel variants (CUDA-core vs mma.sync) benched side by side, eval chamfer stress.
TAG=${1:-r01p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 540 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -60 | tee $OUT/pytest_gpu.txt
for mode in 0 2 6 14; do
  echo "== bench GNBV_CONV2_TC=$mode"
  GNBV_CONV2_TC=$mode timeout 240 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline 2>$OUT/bench_err_$mode.txt | tee $OUT/bench_mode$mode.json | cut -c1-400
  tail -2 $OUT/bench_err_$mode.txt
done
echo "== chamfer stress (BASELINE configs[4])"
timeout 420 python scripts/chamfer_stress.py --out $OUT/chamfer_stress.json 2>&1 | tail -3
ls -la $OUT

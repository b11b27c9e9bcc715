#!/bin/bash
# One GPU session: parity tests, smoke, bench (both arms), ncu launch list + full capture of the top kernels.
# Usage (from the repo root, under gpurun):  [SKIP_TESTS=1] [SKIP_NCU=1] bash scripts/gpu_round.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
  echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee $OUT/smoke.txt
fi
echo "== bench native"; timeout 900 python bench.py --gpus 1 --steps 50 --warmup 5 2>$OUT/bench_err.txt | tee $OUT/bench.json
tail -5 $OUT/bench_err.txt
echo "== bench reference"; timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 2>>$OUT/bench_err.txt | tee $OUT/bench_reference.json
if [ -z "$SKIP_NCU" ]; then
  echo "== ncu launch list"
  timeout 900 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file $OUT/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1
  echo "== ncu full"
  timeout 1200 ncu --nvtx --nvtx-include "timed/" --set full --clock-control none --import-source on \
      -k regex:'grid_update_kernel|scan_raycast_kernel|conv2_fwd_kernel|conv2_wgrad_kernel|conv2_dgrad_kernel|conv1_wgrad_tma_kernel|conv1_fwd_tma_kernel' \
      -c 14 -o $OUT/prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_bench.log 2>&1
fi
ls -la $OUT

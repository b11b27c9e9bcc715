#!/bin/bash
# One GPU session: parity tests, smoke, bench (both arms), ncu launch list + full capture of the top kernels.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_round.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
[ -n "$SKIP_TESTS" ] || timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench native"; timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 2>$OUT/bench_err.txt | tee $OUT/bench.json
tail -5 $OUT/bench_err.txt
echo "== bench reference"; timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 2>>$OUT/bench_err.txt | tee $OUT/bench_reference.json
echo "== ncu launch list"
timeout 600 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'grid_update_kernel|scan_raycast_kernel' -s 6 -c 4 \
    -o $OUT/prof python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_bench.log 2>&1
ls -la $OUT

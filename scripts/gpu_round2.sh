#!/bin/bash
# One GPU session of round 2: parity suite, smoke, bench.   Usage (repo root, under gpurun): bash scripts/gpu_round2.sh <tag> [quick]
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -40 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 200 python __graft_entry__.py smoke 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== bench native"; timeout 400 python bench.py --gpus 1 --steps 30 --warmup 5 2>$OUT/bench_err.txt | tee $OUT/bench.json | cut -c1-300
tail -5 $OUT/bench_err.txt
ls $OUT

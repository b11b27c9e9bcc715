#!/bin/bash
# Final GPU session of the round: full parity suite, smoke, bench (both arms + kernel-variant runs), ncu launch list,
# one PPO iteration (BASELINE configs[2]), eval chamfer stress (configs[4]).  Ordered by importance.
TAG=${1:-r01v}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 330 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -25 | tee $OUT/pytest_gpu.txt
echo "== bench native"; timeout 120 python bench.py --gpus 1 --steps 30 --warmup 5 2>$OUT/bench_err.txt | tee $OUT/bench.json | cut -c1-200
echo "== bench reference"; timeout 150 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 2>>$OUT/bench_err.txt | tee $OUT/bench_reference.json | cut -c1-300
run() { echo "== bench $1"; env $1 timeout 100 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline 2>>$OUT/bench_err.txt | tee $OUT/bench_$2.json | cut -c1-200; }
run "GNBV_CONV2_TC=62" c62
run "GNBV_CONV2_TC=62 GNBV_GEMM_MMA=1" c62_gemm1
echo "== ppo tests with the tensor-core GEMM"; GNBV_GEMM_MMA=1 GNBV_CONV2_TC=62 timeout 100 python -m pytest tests/test_ppo_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -3 | tee $OUT/pytest_ppo_gemm1.txt
echo "== smoke"; timeout 100 python __graft_entry__.py smoke 2>&1 | tail -2 | tee $OUT/smoke.txt
echo "== ncu launch list"
timeout 150 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1
echo "== ppo iteration"; PPO_ITER_FAST=1 timeout 150 python scripts/ppo_iteration.py 2>&1 | tail -1 | tee $OUT/ppo_iteration.json | cut -c1-400
echo "== chamfer stress"; timeout 100 python scripts/chamfer_stress.py --out $OUT/chamfer_stress.json 2>&1 | tail -2 | cut -c1-300
ls $OUT

#!/bin/bash
# One GPU session: full parity suite, smoke, bench (both arms + kernel-variant runs), ncu launch list (+ optional --set full
# capture of the tensor-core kernels), one PPO iteration (BASELINE configs[2]), eval chamfer stress (configs[4]).  Ordered by
# importance.   Usage (repo root, under gpurun):  [WITH_NCU_FULL=1] bash scripts/gpu_round.sh <tag>
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 330 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -25 | tee $OUT/pytest_gpu.txt
echo "== bench native"; timeout 120 python bench.py --gpus 1 --steps 30 --warmup 5 2>$OUT/bench_err.txt | tee $OUT/bench.json | cut -c1-200
echo "== bench reference"; timeout 150 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 2>>$OUT/bench_err.txt | tee $OUT/bench_reference.json | cut -c1-300
run() { echo "== bench $1"; env $1 timeout 100 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline 2>>$OUT/bench_err.txt | tee $OUT/bench_$2.json | cut -c1-200; }
run "GNBV_CONV2_TC=0 GNBV_CONV1_MMA=0 GNBV_GEMM_MMA=0" cuda_cores
run "GNBV_GEMM_MMA=0" fp32_gemm
echo "== smoke"; timeout 100 python __graft_entry__.py smoke 2>&1 | tail -2 | tee $OUT/smoke.txt
echo "== ncu launch list"
timeout 150 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1
if [ -n "$WITH_NCU_FULL" ]; then
  timeout 400 ncu --nvtx --nvtx-include "timed/" --set full --clock-control none --import-source on \
      -k regex:"mma_kernel|wgrad_staged|grid_update|scan_raycast" -c 12 -o $OUT/prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
fi
echo "== ppo iteration"; PPO_ITER_FAST=1 timeout 150 python scripts/ppo_iteration.py 2>&1 | tail -1 | tee $OUT/ppo_iteration.json | cut -c1-400
echo "== chamfer stress"; timeout 100 python scripts/chamfer_stress.py --out $OUT/chamfer_stress.json 2>&1 | tail -2 | cut -c1-300
ls $OUT

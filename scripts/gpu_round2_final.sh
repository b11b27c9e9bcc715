#!/bin/bash
# Round-2 evidence session (one GPU): full parity suite, smoke, bench (both arms), ncu launch list + --set full of the hot kernels,
# eval chamfer stress.   Usage (repo root, under gpurun): bash scripts/gpu_round2_final.sh <tag>
TAG=${1:-r02z}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -30 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 200 python __graft_entry__.py smoke 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== bench native"; timeout 500 python bench.py --gpus 1 --steps 30 --warmup 5 2>$OUT/bench_err.txt | grep "^{" | tee $OUT/bench.json | cut -c1-300
echo "== bench reference"; timeout 400 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 2>>$OUT/bench_err.txt | grep "^{" | tee $OUT/bench_reference.json | cut -c1-300
echo "== ncu launch list"
timeout 200 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ppo > $OUT/ncu_launch_bench.log 2>&1
echo "== ncu full"
timeout 400 ncu --nvtx --nvtx-include "timed/" --set full --clock-control none --import-source on \
    -k regex:"scan_raycast|grid_update_sparse|conv1_fwd_mma|conv2_fwd_mma|tc_gemm|wgrad_staged|conv2_dgrad_mma|conv1_wgrad_mma" -c 10 \
    -o $OUT/prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ppo > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log | cut -c1-200
echo "== tcgen05 micro-benchmarks"; (timeout 60 ./scripts/micro/mma_rate; timeout 60 ./scripts/micro/handoff_latency) > $OUT/micro_tcgen05.txt 2>&1; tail -4 $OUT/micro_tcgen05.txt
echo "== chamfer stress"; timeout 150 python scripts/chamfer_stress.py --out $OUT/chamfer_stress.json 2>&1 | tail -1 | cut -c1-300
ls -la $OUT

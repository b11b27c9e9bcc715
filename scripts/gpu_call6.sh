#!/bin/bash
TAG=${1:-r01u}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest"; timeout 700 python -m pytest tests/test_policy_gpu.py tests/test_ppo_gpu.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -30 | tee $OUT/pytest_gpu.txt
run() { echo "== bench $1"; env $1 timeout 240 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline 2>$OUT/bench_err.txt | tee $OUT/bench_$2.json | cut -c1-240; }
run "GNBV_CONV2_TC=14" c14
run "GNBV_CONV2_TC=30" c30
run "GNBV_CONV2_TC=30 GNBV_GEMM_MMA=1" c30_gemm1
GNBV_CONV2_TC=30 GNBV_GEMM_MMA=1 timeout 400 ncu --nvtx --nvtx-include "timed/" --set full --clock-control none --import-source on \
    -k regex:"wgrad_staged|dgrad_mma|sgemm_mma" -c 8 -o $OUT/prof_mma python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log | cut -c1-200
ls $OUT

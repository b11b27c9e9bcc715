#!/bin/bash
OUT=gpurun_out/r01w; mkdir -p $OUT
GNBV_GEMM_MMA=1 timeout 120 python -m pytest tests/test_policy_gpu.py tests/test_ppo_gpu.py tests/test_evaluation.py -q -m gpu -p no:cacheprovider -k "not variant and not tcgen05" 2>&1 | tail -6 | tee $OUT/pytest_gemm1.txt
GNBV_GEMM_MMA=1 timeout 60 python __graft_entry__.py smoke 2>&1 | tail -1 | tee $OUT/smoke_gemm1.txt

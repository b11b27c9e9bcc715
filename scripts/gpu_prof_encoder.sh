#!/bin/bash
TAG=${1:-r01c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_policy_gpu.py tests/test_ppo_gpu.py -q --tb=line 2>&1 | tail -5 | tee $OUT/pytest.txt
timeout 300 python scripts/time_encoder.py 64 2>&1 | tail -4 | tee $OUT/time_encoder.txt
timeout 300 python scripts/time_encoder.py 20 2>&1 | tail -4 | tee -a $OUT/time_encoder.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file $OUT/launches_enc.csv \
    python scripts/time_encoder.py 64 > $OUT/ncu_enc.log 2>&1
ls -la $OUT

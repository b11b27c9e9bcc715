#!/bin/bash
OUT=gpurun_out/r01x; mkdir -p $OUT
timeout 100 python bench.py --gpus 1 --steps 30 --warmup 5 2>$OUT/bench_err.txt | tee $OUT/bench.json | cut -c1-200
tail -2 $OUT/bench_err.txt

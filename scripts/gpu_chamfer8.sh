#!/bin/bash
# BASELINE configs[4] on 8 GPUs: one independent process per GPU (the eval accuracy path has no cross-env exchange), 256 envs each.
# The host has 16 cores: cap the CPU threads per process (8 x 16 OpenMP threads spinning on 16 cores stalled the first attempt).
TAG=${1:-r02c8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export OMP_NUM_THREADS=2 MKL_NUM_THREADS=2
nproc > $OUT/host_cores.txt
t0=$(date +%s.%N)
for r in 0 1 2 3 4 5 6 7; do CUDA_VISIBLE_DEVICES=$r timeout 150 python scripts/chamfer_stress.py --out $OUT/chamfer_stress_8gpu_rank$r.json > $OUT/chamfer_rank$r.log 2>&1 & done
wait
t1=$(date +%s.%N)
python - <<PY
import json, glob
rs = [json.load(open(f)) for f in sorted(glob.glob("$OUT/chamfer_stress_8gpu_rank*.json"))]
out = {"ranks": len(rs), "envs_total": 256 * len(rs), "wall_s_all_ranks_concurrent": round($t1 - $t0, 2),
       "chamfer_grid_ms_all_envs_per_rank": [round(r["chamfer_grid_ms_all_envs"], 3) for r in rs],
       "dedup_decode_ms_per_rank": [round(r["dedup_decode_ms_all_envs"], 3) for r in rs],
       "eval_env_step_ms_median_per_rank": [round(r["eval_env_step_ms"]["median"], 3) for r in rs]}
json.dump(out, open("$OUT/chamfer_stress_8gpu_summary.json", "w"), indent=1)
print(json.dumps(out))
PY
tail -2 $OUT/chamfer_rank0.log | cut -c1-200

#!/bin/bash
# Two-GPU session (gpurun --gpus 2): the NCCL tests of PPO_Grid_Obs.train(), the bench at N = 2 as the driver launches it, and the eval
# accuracy stress as two concurrent independent processes (one per GPU).
TAG=${1:-r02t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest multi-gpu"; timeout 400 python -m pytest tests/test_ppo_multi_gpu.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -5 | tee $OUT/pytest_multi_gpu.txt
echo "== bench N=2"
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 30 --warmup 5 2> $OUT/bench_2gpu.err | grep "^{" > $OUT/bench_2gpu.json
python -c "import sys,json; d=json.loads(open('$OUT/bench_2gpu.json').read()); p=d.get('ppo_iteration') or {}; print('N', d['n_gpus'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'ppo env-steps/s', round(p.get('env_steps_per_sec',0)), 'ms/update', round(p.get('ms_per_minibatch_update',0),3))" || tail -5 $OUT/bench_2gpu.err
echo "== chamfer stress, 2 concurrent processes"
export OMP_NUM_THREADS=4 MKL_NUM_THREADS=4
t0=$(date +%s.%N)
for r in 0 1; do CUDA_VISIBLE_DEVICES=$r timeout 150 python scripts/chamfer_stress.py --out $OUT/chamfer_stress_2gpu_rank$r.json > $OUT/chamfer_rank$r.log 2>&1 & done
wait
t1=$(date +%s.%N)
python - <<PY
import json, glob
rs = [json.load(open(f)) for f in sorted(glob.glob("$OUT/chamfer_stress_2gpu_rank*.json"))]
out = {"ranks": len(rs), "envs_total": 256 * len(rs), "wall_s_all_ranks_concurrent": round($t1 - $t0, 2),
       "chamfer_grid_ms_all_envs_per_rank": [round(r["chamfer_grid_ms_all_envs"], 3) for r in rs],
       "dedup_decode_ms_per_rank": [round(r["dedup_decode_ms_all_envs"], 3) for r in rs],
       "eval_env_step_ms_median_per_rank": [round(r["eval_env_step_ms"]["median"], 3) for r in rs]}
json.dump(out, open("$OUT/chamfer_stress_2gpu_summary.json", "w"), indent=1)
print(json.dumps(out))
PY
ls $OUT

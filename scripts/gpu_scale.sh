#!/bin/bash
# Weak-scaling session (run with gpurun --gpus 8): bench at N = 1, 2, 4, 8 (as the driver launches it) + eval chamfer stress on 1 and 8 GPUs.
TAG=${1:-r02s}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for N in 1 2 4 8; do
  echo "== bench N=$N"
  if [ $N -eq 1 ]; then
    timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_${N}gpu.json 2> $OUT/bench_${N}gpu.err
  else
    timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_${N}gpu.json 2> $OUT/bench_${N}gpu.err
  fi
  grep "^{" $OUT/bench_${N}gpu.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); p=d.get('ppo_iteration') or {}; print('N', d['n_gpus'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'ppo env-steps/s', round(p.get('env_steps_per_sec',0)), 'ms/update', round(p.get('ms_per_minibatch_update',0),3))" || tail -5 $OUT/bench_${N}gpu.err
done
echo "== chamfer stress 1 GPU"; timeout 200 python scripts/chamfer_stress.py --out $OUT/chamfer_stress_1gpu.json 2>&1 | tail -1 | cut -c1-200
echo "== chamfer stress 8 GPUs (one process per GPU, 256 envs each)"
for r in 0 1 2 3 4 5 6 7; do CUDA_VISIBLE_DEVICES=$r timeout 300 python scripts/chamfer_stress.py --out $OUT/chamfer_stress_8gpu_rank$r.json > $OUT/chamfer_rank$r.log 2>&1 & done
wait
python - <<PY
import json, glob
rs = [json.load(open(f)) for f in sorted(glob.glob("$OUT/chamfer_stress_8gpu_rank*.json"))]
print("ranks", len(rs), [round(r.get("chamfer_grid_ms_all_envs", r.get("chamfer_ms", 0)), 2) for r in rs])
PY
ls $OUT

GNBV_CONV2_TC=126 timeout 120 python scripts/diag_grads.py 64 8 2>&1 | grep -v Warn | sed -n 2,4p
GNBV_CONV2_TC=126 timeout 120 python scripts/diag_grads.py 20 9 2>&1 | grep -v Warn | sed -n 2,4p
for d in 32; do echo "== dbg $d"; GNBV_TS_DEBUG=$d GNBV_CONV2_TC=126 timeout 100 python scripts/micro/ts_prof.py 2>&1 | grep -v Warn; done
for d in 0; do
GNBV_TS_DEBUG=$d GNBV_CONV2_TC=126 timeout 100 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-ppo 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('dbg', $d, 'fwd.conv2 ms', round(d['encoder_kernel_ms']['fwd.conv2'],4), 'bn2', round(d['encoder_kernel_ms']['fwd.bn2_stats+apply'],4))"
done

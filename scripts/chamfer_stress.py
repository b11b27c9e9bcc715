"""BASELINE configs[4] -- eval accuracy stress: E envs x 100 k GT surface points x a 50-view scanned-point history.

Runs the real pipeline: `Env_Eval_GenNBV` (synthetic 128x128 sensor) is stepped for 50 look-at views, so every env's
history holds up to 50 x 16384 lattice keys; then, for all envs at once, dedup (torch.unique on the packed keys) + decode +
exact bidirectional 1-NN chamfer against the GT clouds, timed with CUDA events.  The grid search is compared with the
P1*P2 scan kernel on a few envs (bit-identical per-point minima, and the speed-up).

    python scripts/chamfer_stress.py [--envs 256] [--gt 100000] [--views 50] [--brute-envs 2] [--out file.json]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gennbv_b200 import chamfer, synth  # noqa: E402
from gennbv_b200.config import Config_GenNBV_Eval  # noqa: E402
from gennbv_b200.env_eval import Env_Eval_GenNBV  # noqa: E402
from gennbv_b200.sensors import SyntheticHouseSensor  # noqa: E402


def ev_time(fn, reps=3):
    best = None
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        best = ms if best is None else min(best, ms)
    return best, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=256)
    ap.add_argument("--gt", type=int, default=100000)
    ap.add_argument("--views", type=int, default=50)
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--grid", type=int, default=64)
    ap.add_argument("--brute-envs", type=int, default=2)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    dev = "cuda:0"
    E, H = a.envs, a.size

    class Cfg(Config_GenNBV_Eval):
        max_episode_length = a.views + 1            # no time-out inside the measured history

    scenes = synth.make_house_scenes(8, a.grid, seed=0)
    pc_gt = synth.gt_point_clouds(scenes.params, E, a.gt, seed=0)
    env = Env_Eval_GenNBV(Cfg(), sim_device=dev, sensor=SyntheticHouseSensor(scenes.params, H, H, device=dev, with_rgb=False),
                          grid_gt=scenes.grid_gt, pc_gt=pc_gt, num_envs=E)
    gen = torch.Generator().manual_seed(1)
    env.reset()
    step_ms = []
    for v in range(a.views):
        act = synth.sample_lookat_actions(scenes.params, E, gen).to(dev)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s.record()
        env.step(act)
        e.record()
        torch.cuda.synchronize()
        step_ms.append(s.elapsed_time(e))
    counts = env._pts_count.tolist()
    assert int(env._pts_overflow) == 0

    dedup_ms, (pts, sizes) = ev_time(lambda: env.dedup_clouds(list(range(E))), reps=2)
    offs = [0]
    for n in sizes:
        offs.append(offs[-1] + n)
    clouds = [pts[offs[i]:offs[i + 1]] for i in range(E)]
    gts = env.pc_gt
    gt_packed, gt_sizes = torch.cat(gts, 0).contiguous(), [int(g.shape[0]) for g in gts]
    C = chamfer.default_cells_per_axis(max(max(sizes), a.gt))
    grid_ms, (cx, cy) = ev_time(lambda: chamfer.chamfer_terms_packed(pts, sizes, gt_packed, gt_sizes, method="grid"), reps=3)
    nb = min(a.brute_envs, E)
    res = {
        "workload": f"{E} envs x {a.gt} GT points x {a.views}-view history ({H}x{H} depth)",
        "history_keys_per_env": {"min": min(counts), "mean": sum(counts) / E, "max": max(counts)},
        "dedup_points_per_env": {"min": min(sizes), "mean": sum(sizes) / E, "max": max(sizes)},
        "eval_env_step_ms": {"median": sorted(step_ms)[len(step_ms) // 2], "includes": "sensor synthesis + env.step + history append"},
        "dedup_decode_ms_all_envs": dedup_ms, "dedup_how": "gnbv_pack_env_keys + in-tree radix sort / unique (gnbv_sort_unique_u64) + one decode launch",
        "chamfer_grid_ms_all_envs": grid_ms,
        "cells_per_axis": C,
        "accuracy": {"min": float((cx + cy).min()), "mean": float((cx + cy).mean()), "max": float((cx + cy).max())},
    }
    # algorithmic bytes of the exact-NN formulation: (P1 + P2) * (12 B read + 4 B written) per env
    alg_bytes = sum((s + a.gt) * 16 for s in sizes)
    res["chamfer_grid_algorithmic_GBps"] = alg_bytes / (grid_ms * 1e-3) / 1e9
    res["pairs_per_env_mean"] = sum(s * a.gt for s in sizes) / E
    if nb > 0:
        xs, ys = clouds[:nb], gts[:nb]
        brute_ms, b = ev_time(lambda: chamfer.chamfer_terms(xs, ys, method="brute", return_min=True), reps=1)
        sub_ms, g = ev_time(lambda: chamfer.chamfer_terms(xs, ys, method="grid", return_min=True), reps=2)
        res["brute_vs_grid"] = {
            "envs": nb, "brute_ms": brute_ms, "grid_ms": sub_ms, "speedup": brute_ms / sub_ms,
            "brute_pair_rate_Tpairs_s": 2 * sum(s * a.gt for s in sizes[:nb]) / (brute_ms * 1e-3) / 1e12,
            "minima_bit_identical": bool(torch.equal(b[2], g[2]) and torch.equal(b[3], g[3])),
            "terms_bit_identical": bool(torch.equal(b[0], g[0]) and torch.equal(b[1], g[1])),
        }
    line = json.dumps(res)
    print(line)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with open(a.out, "w") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    t0 = time.time()
    main()
    print(f"# wall {time.time() - t0:.1f} s", file=sys.stderr)

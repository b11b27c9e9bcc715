"""Shared test helpers: golden-fixture loading and the replay harness used by both the
oracle tests (CPU) and the CUDA parity tests (GPU)."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ENV_GOLDENS = ["env_g20", "env_g64", "env_g20_long"]


class EnvGolden:
    """tests/golden/env_*.npz (written by oracle/gen_golden.py from the reference's own env)."""

    def __init__(self, name):
        d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.d = d
        self.N, self.H, self.W, self.G, self.S, self.T, self.max_len, self.seed = (int(v) for v in d["meta"])
        G, S = self.G, self.S
        gt = np.unpackbits(d["grid_gt_file"])[: S * G ** 3].reshape(S, G, G, G).astype(np.float32)
        self.grid_gt = gt[np.arange(self.N) % S]                       # env_train_gennbv.py:86-92
        self.prob = d["prob_pal"][d["prob_codes"]].view(np.float32)     # [T+1,N,G,G,G]
        self.scan = d["scan_pal"][d["scan_codes"]].view(np.float32)
        self.tri = d["tri"].astype(np.float32)

    def __getattr__(self, k):
        return self.d[k]


def replay_voxelize(g, step_fn):
    """Drives `step_fn(depth_raw, seg, c2w, pose_xyz, prob, scan) -> (tri, cov_sum)` (in-place on prob / scan,
    numpy fp32) through the golden roll-out and checks every step against the reference's recorded state.
    Returns the number of compared steps."""
    N, G = g.N, g.G
    prob = np.zeros((N, G, G, G), np.float32)
    scan = np.zeros((N, G, G, G), np.float32)
    for t in range(g.T + 1):
        tri, cov = step_fn(g.depth[t], g.seg[t], g.c2w[t], np.ascontiguousarray(g.poses[t][:, :3]), prob, scan)
        np.testing.assert_array_equal(tri, g.tri[t], err_msg=f"tri-class grid, step {t}")
        done = g.done[t].astype(bool)
        ratio = cov / g.num_valid_voxel_gt
        np.testing.assert_array_equal(ratio[~done], g.ratio[t][~done], err_msg=f"coverage ratio, step {t}")
        prob[done] = 0.0      # reset_idx (env_train_gennbv.py:413-417)
        scan[done] = 0.0
        np.testing.assert_array_equal(prob, g.prob[t], err_msg=f"prob_grid, step {t}")
        np.testing.assert_array_equal(scan, g.scan[t], err_msg=f"scanned_gt_grid, step {t}")
    return g.T + 1

"""CPU: the oracle's restatement of the eval env's point history (back-projection -> 1 cm rounding -> unique) and of the
chamfer accuracy reproduces what the reference's own Env_Eval_GenNBV produced (tests/golden/env_eval_g20.npz, recorded by
oracle/gen_golden.py --eval), including the history clears at episode end and inside reset()."""
import os

import numpy as np
import torch

import oracle as c_oracle
from gennbv_b200 import synth
from helpers import GOLDEN_DIR


def load():
    return np.load(os.path.join(GOLDEN_DIR, "env_eval_g20.npz"))


def test_point_history_and_accuracy_restatement_match_the_reference_env():
    d = load()
    N = int(d["meta"][0])
    kinv = d["inv_intri"]
    origins = torch.from_numpy(d["env_origins"])
    hist = [np.zeros((0, 3), np.float32) for _ in range(N)]
    cloud_off = np.concatenate([[0], np.cumsum(d["cloud_sizes"])])
    ci = 0
    checked_acc = 0
    seen = {}
    for call in range(d["rew"].shape[0]):
        if d["is_reset"][call]:
            hist = [np.zeros((0, 3), np.float32) for _ in range(N)]          # reset_idx(all) before the observation pass
        c2w = synth.c2w_from_view_matrix(torch.from_numpy(d["view"][call]), origins).numpy()
        depth = c_oracle.post_process_depth(d["depth"][call])
        world, fg = c_oracle.back_projection(depth, d["seg"][call], kinv, c2w)
        for e in range(N):
            hist[e] = np.concatenate([hist[e], world[e][fg[e]]], 0)
        sizes_before_clear = np.array([h.shape[0] for h in hist])
        for e in np.nonzero(d["done"][call])[0]:
            if ci < len(d["cloud_call"]) and d["cloud_call"][ci] == call:
                want = d["cloud_points"][cloud_off[ci]:cloud_off[ci + 1]]
                got = c_oracle.round_1cm_unique(hist[e])
                np.testing.assert_array_equal(got, want, err_msg=f"dedup cloud, call {call}, env {e}")
                cx, cy = c_oracle.chamfer(got, d["pc_gt"][e])
                if str(e) not in seen:
                    seen[str(e)] = cx + cy
                ci += 1
            hist[e] = np.zeros((0, 3), np.float32)
        if d["is_reset"][call]:
            acc_now, seen = dict(seen), {}                                  # reset() returns the old dict and rebinds a new one
        else:
            acc_now = seen
        for e in range(N):
            want = d["acc"][call][e]
            if np.isnan(want):
                assert str(e) not in acc_now
            else:
                assert abs(acc_now[str(e)] - want) <= 1e-5 * want
                checked_acc += 1
        np.testing.assert_array_equal(np.array([h.shape[0] for h in hist]), d["hist_sizes"][call])
        assert (sizes_before_clear >= d["hist_sizes"][call]).all()
    assert ci == len(d["cloud_call"]) and checked_acc > 0


def test_rounding_is_half_even_in_fp32_and_rows_come_out_sorted():
    p = np.array([[0.125, -0.125, 0.135], [0.005, 0.015, 0.025], [0.125, -0.125, 0.135], [-0.004, 0.0, 2.0]], np.float32)
    r = c_oracle.round_1cm_unique(p)
    want = torch.unique(torch.round(torch.from_numpy(p), decimals=2), dim=0).numpy()
    np.testing.assert_array_equal(r, want)
    assert r.shape[0] == 3


def test_chamfer_oracle_equals_the_brute_force_definition():
    """pytorch3d.loss.chamfer_distance defaults (norm 2, point_reduction "mean", single cloud pair): mean over x of the squared
    distance to the nearest y, plus the same with the roles swapped.  The oracle's k-d tree evaluation against all P1 x P2 pairs."""
    rng = np.random.default_rng(7)
    for p1, p2 in ((1, 1), (17, 5), (300, 411)):
        x = rng.normal(size=(p1, 3)).astype(np.float32)
        y = (rng.normal(size=(p2, 3)) * 1.5 + 0.3).astype(np.float32)
        d2 = ((x[:, None, :].astype(np.float64) - y[None, :, :].astype(np.float64)) ** 2).sum(-1)
        want = d2.min(1).mean() + d2.min(0).mean()
        got = c_oracle.chamfer(x, y)
        got = float(got[0] + got[1]) if isinstance(got, (tuple, list)) else float(got)
        assert abs(got - want) <= 1e-12 * max(1.0, want), (p1, p2, got, want)

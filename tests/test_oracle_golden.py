"""CPU: the C oracle (oracle/gennbv_oracle.c) against the fixtures recorded from the reference's own env."""
import numpy as np
import pytest

import oracle as c_oracle
from helpers import ENV_GOLDENS, EnvGolden, replay_voxelize


@pytest.mark.parametrize("name", ENV_GOLDENS)
def test_voxelize_step_matches_reference_rollout(name):
    g = EnvGolden(name)

    def step(depth_raw, seg, c2w, pose_xyz, prob, scan):
        out = c_oracle.voxelize_step(depth_raw, seg, g.inv_intri, c2w, g.range_gt, g.voxel_size_gt, pose_xyz,
                                     g.grid_gt, prob, scan, raw_depth=True)
        return out["tri"], out["cov_sum"]

    assert replay_voxelize(g, step) == g.T + 1


@pytest.mark.parametrize("name", ENV_GOLDENS)
def test_depth_post_process(name):
    g = EnvGolden(name)
    d = c_oracle.post_process_depth(g.depth[0])
    assert np.isfinite(d).all() and (d >= 0).all() and (d <= 50).all()
    assert ((g.depth[0] == -np.inf) == (d == 0)).all()


def test_goldens_cover_resets_and_terminations():
    g = EnvGolden("env_g20")
    assert g.done[1:].sum() >= 8 and g.time_out[1:].any()
    # after a done step the env forces init_action (env_train_gennbv.py:247-253)
    t = int(np.argmax(g.done[1:, 0])) + 1
    assert (g.applied_actions[t + 1][0] == np.array([40, 40, 50, 0, 12, 0])).all()
    gl = EnvGolden("env_g20_long")
    assert gl.done[1:].sum() >= 1 and not gl.time_out.any()      # coverage > 0.99 termination (ratio is zeroed by the reset)
    # fp32 sequential decrement: 0.49999988 (= 1 - 10 x 0.05f) appears, i.e. occupied -> unknown
    assert np.float32(0.49999988) in gl.prob


@pytest.mark.parametrize("name", ENV_GOLDENS)
def test_torch_restatement_matches_reference_rollout(name):
    """oracle/torch_ref.py (the `--impl reference` arm of bench.py) reproduces the recorded reference states."""
    import torch
    import torch_ref
    g = EnvGolden(name)
    pix = torch_ref.pixel_grid(g.H, g.W)
    T = torch.from_numpy

    def step(depth_raw, seg, c2w, pose_xyz, prob, scan):
        p, s = T(prob), T(scan)      # share memory: in-place updates land in the numpy state
        tri, cov = torch_ref.voxelize_step(T(depth_raw), T(seg), T(g.inv_intri), T(c2w), T(g.range_gt),
                                           T(g.voxel_size_gt), T(pose_xyz), T(g.grid_gt), p, s, pix)
        return tri.numpy(), cov.numpy()

    assert replay_voxelize(g, step) == g.T + 1

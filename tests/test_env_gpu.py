"""GPU: the drop-in `Env_Train_GenNBV` replayed against the roll-outs recorded from the reference's own env
(same raw sensor frames and Isaac-convention view matrices): every returned tensor and every piece of state the
reference exposes is compared bit for bit, across resets, time-outs, forced init actions and coverage termination."""
import numpy as np
import pytest
import torch

from gennbv_b200.config import Config_GenNBV_Train
from gennbv_b200.env import Env_Train_GenNBV
from gennbv_b200.sensors import ReplaySensor
from gennbv_b200.wrapper import EnvWrapperGenNBVTrain
from helpers import ENV_GOLDENS, EnvGolden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def make_env(g, with_rgb=True):
    class Cfg(Config_GenNBV_Train):
        max_episode_length = g.max_len

        class rewards(Config_GenNBV_Train.rewards):
            only_positive_rewards = False          # train_gennbv.py:101-106

    S, G = g.S, g.G
    gt = np.unpackbits(g.grid_gt_file)[: S * G ** 3].reshape(S, G, G, G).astype(np.float32)
    lo, hi = g.grid_centres_lohi[:, 0], g.grid_centres_lohi[:, 1]
    grid = torch.zeros(S, G, G, G, 4)
    for s in range(S):       # the file layout: voxel centres + occupancy (only the corner differences are consumed)
        ax = [torch.linspace(float(lo[s, a]), float(hi[s, a]), G) for a in range(3)]
        cx, cy, cz = torch.meshgrid(*ax, indexing="ij")
        grid[s, ..., 0], grid[s, ..., 1], grid[s, ..., 2] = cx, cy, cz
    grid[..., 3] = torch.from_numpy(gt)
    sensor = ReplaySensor(g.depth, g.seg, g.rgb if with_rgb else None, g.view, DEV)
    env = Env_Train_GenNBV(Cfg(), sim_device=DEV, sensor=sensor, grid_gt=grid, num_envs=g.N)
    # GT metadata must be the reference's own fp32 values (derived from the real file's centres)
    for name, ref in (("range_gt", g.range_gt), ("voxel_size_gt", g.voxel_size_gt), ("num_valid_voxel_gt", g.num_valid_voxel_gt)):
        getattr(env, name).copy_(torch.from_numpy(ref))
    np.testing.assert_array_equal(env.inv_intri.cpu().numpy(), g.inv_intri)
    np.testing.assert_array_equal(env.env_origins.numpy(), g.env_origins)
    np.testing.assert_allclose([env.reward_scales[k] for k in ("surface_coverage", "short_path", "termination")],
                               g.reward_scales, rtol=0, atol=0)
    return env


def check(env, g, t, obs, rew=None, done=None, infos=None):
    eq = np.testing.assert_array_equal
    c = lambda x: x.detach().cpu().numpy()
    eq(c(obs["state"]), g.state[t], err_msg=f"obs.state step {t}")
    eq(c(obs["state_rgb"]), g.state_rgb[t], err_msg=f"obs.state_rgb step {t}")
    eq(c(obs["grid"]), g.tri[t], err_msg=f"obs.grid step {t}")
    eq(c(env.prob_grid), g.prob[t], err_msg=f"prob_grid step {t}")
    eq(c(env.scanned_gt_grid), g.scan[t], err_msg=f"scanned_gt_grid step {t}")
    eq(c(env.reward_ratio_buf[-1]), g.ratio[t], err_msg=f"ratio step {t}")
    eq(c(env.episode_length_buf), g.ep_len[t], err_msg=f"episode_length_buf step {t}")
    eq(c(env.extras["time_outs"]), g.time_out[t].astype(bool), err_msg=f"infos.time_outs step {t}")
    assert not c(env.reset_buf).any()
    if rew is not None:
        eq(c(rew), g.rew[t], err_msg=f"reward step {t}")
        eq(c(done), g.done[t].astype(bool), err_msg=f"done step {t}")
        assert infos is env.extras and set(infos["episode"]) >= {"rew_surface_coverage", "episode_reward", "episode_length"}


@pytest.mark.parametrize("name", ENV_GOLDENS)
def test_env_rollout_matches_reference(name):
    g = EnvGolden(name)
    env = make_env(g)
    obs = env.reset()
    check(env, g, 0, obs)
    prev_obs = prev_done = None
    for t in range(g.T):
        a = torch.from_numpy(g.actions[t]).to(DEV)
        obs, rew, done, infos = env.step(a)
        check(env, g, t + 1, obs, rew, done, infos)
        if prev_obs is not None:     # the previous step's tensors stay valid for one more step (SB3 keeps them)
            np.testing.assert_array_equal(prev_obs[0].cpu().numpy(), g.tri[t])
            np.testing.assert_array_equal(prev_done.cpu().numpy(), g.done[t].astype(bool))
        prev_obs, prev_done = (obs["grid"],), done


def test_episode_statistics_match_reference_semantics():
    g = EnvGolden("env_g20")
    env = make_env(g)
    env.reset()
    ep_rewards, cur = [], g.rew[0].copy()        # reset() itself runs one reward pass (env_train_gennbv.py:243)
    for t in range(g.T):
        _, rew, done, infos = env.step(torch.from_numpy(g.actions[t]).to(DEV))
        cur += g.rew[t + 1]
        for n in np.nonzero(g.done[t + 1])[0]:
            ep_rewards.append(float(cur[n])); cur[n] = 0
        want = np.mean(ep_rewards[-100:]) if ep_rewards else 0.0
        assert abs(float(infos["episode"]["episode_reward"]) - want) < 1e-5
    assert float(infos["episode"]["episode_length"]) == g.max_len


def test_wrapper_returns_flat_layout_without_copy():
    g = EnvGolden("env_g20")
    env = EnvWrapperGenNBVTrain(make_env(g))
    flat = env.reset()
    N, G = g.N, g.G
    assert flat.shape == (N, 600 + G ** 3 + 2 * 64 * 64) and flat.data_ptr() == env.obs_flat.data_ptr()
    assert env.observation_space.shape == (flat.shape[1],) and env.action_space.nvec.tolist() == [81, 81, 51, 1, 13, 13]
    want = np.concatenate([g.state[0].reshape(N, -1), g.tri[0].reshape(N, -1), g.state_rgb[0].reshape(N, -1)], 1)
    np.testing.assert_array_equal(flat.cpu().numpy(), want)
    env.episode_length_buf = torch.zeros(3)              # lands on the wrapper, not on the env
    assert env._gym_env.episode_length_buf.shape == (N,)

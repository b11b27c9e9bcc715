"""GPU: rollout buffer, collect_rollouts glue and the fused PPO update against a plain-PyTorch re-run of the
reference's train() lines (oracle/encoder_ref.py) with identical weights and data."""
import numpy as np
import pytest
import torch

import encoder_ref
import oracle as c_oracle
from gennbv_b200.buffers import TensorRolloutBuffer_Grid_Obs
from gennbv_b200.ppo import PPO_Grid_Obs
from gennbv_b200.spaces import Box, MultiDiscrete
from gennbv_b200.wrapper import EnvWrapperGenNBVTrain
from helpers import EnvGolden
from test_env_gpu import make_env

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
POLICY_KW = lambda: dict(net_arch=[], features_extractor_kwargs=dict(
    encoder_param={"hidden_shapes": [256, 256], "visual_dim": 256},
    net_param={"transformer_params": [[1, 256], [1, 256]], "append_hidden_shapes": [256, 256]},
    state_input_shape=(600,), visual_input_shape=(100, 48, 48)))


def make_algo(env, n_steps, batch_size, n_epochs, target_kl=None, seed=3):
    algo = PPO_Grid_Obs(env=env, learning_rate=1e-4, n_steps=n_steps, batch_size=batch_size, n_epochs=n_epochs, gamma=0.99,
                        gae_lambda=0.95, clip_range=0.2, clip_range_vf=0.2, ent_coef=0.01, vf_coef=0.8, max_grad_norm=1,
                        target_kl=target_kl, policy_kwargs=POLICY_KW(), seed=seed, device=DEV)
    ref = encoder_ref.PolicyRef(env.grid_size, 600)
    sd = encoder_ref.seeded_state_dict(ref, seed, scale=0.5)
    ref.load_state_dict(sd)
    algo.policy.load_state_dict(sd)
    return algo, ref


def test_collect_rollouts_fills_buffer_like_the_reference_loop():
    g = EnvGolden("env_g20_long")
    env = EnvWrapperGenNBVTrain(make_env(g))
    T = 12
    algo, ref = make_algo(env, n_steps=T, batch_size=9, n_epochs=1)
    algo._setup_learn()
    first_obs = algo._last_obs.clone()
    seen = []

    def spy(loc):
        seen.append(dict(obs=loc["new_obs"].clone(), rew=loc["rewards"].clone(), done=loc["dones"].clone(),
                         tout=loc["infos"]["time_outs"].clone(), actions=loc["actions"].clone(), values=loc["values"].clone()))

    assert algo.collect_rollouts(callback=spy)
    buf = algo.rollout_buffer
    N = g.N
    assert buf.full and algo.num_timesteps == T * N
    np.testing.assert_array_equal(buf.observations[0].cpu().numpy(), first_obs.cpu().numpy())
    for t in range(1, T):          # obs stored at t is the obs returned by step t-1 (the ping-pong buffers keep it alive)
        np.testing.assert_array_equal(buf.observations[t].cpu().numpy(), seen[t - 1]["obs"].cpu().numpy())
        np.testing.assert_array_equal(buf.episode_starts[t, :, 0].cpu().numpy(), seen[t - 1]["done"].cpu().numpy())
    assert buf.episode_starts[0].all()
    ref.eval()
    for t in range(T):
        np.testing.assert_array_equal(buf.actions[t].cpu().numpy(), seen[t]["actions"].float().cpu().numpy())
        # time-out bootstrap: rewards += gamma * V(new_obs)[env 0] * time_outs  (on_policy_algorithm_grid_obs.py:205-208)
        with torch.no_grad():
            v0 = float(ref.value_net(ref.features_extractor(seen[t]["obs"].cpu()))[0, 0])
        want = seen[t]["rew"].cpu() + 0.99 * v0 * seen[t]["tout"].cpu().float()
        np.testing.assert_allclose(buf.rewards[t, :, 0].cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-5)
    # GAE of the stored columns (bit-exact against the oracle)
    with torch.no_grad():
        last_v = algo.policy.predict_values(seen[-1]["obs"]).flatten().cpu().numpy()
    adv, ret = c_oracle.gae(buf.rewards[..., 0].cpu().numpy(), buf.values[..., 0].cpu().numpy(),
                            buf.episode_starts[..., 0].cpu().numpy(), last_v, seen[-1]["done"].cpu().numpy().astype(np.uint8))
    np.testing.assert_array_equal(buf.advantages[..., 0].cpu().numpy(), adv)
    np.testing.assert_array_equal(buf.returns[..., 0].cpu().numpy(), ret)
    # get(): the reference's env-major flat index i = n*T + t under one permutation per rollout
    batch = next(buf.get(9))
    i = buf.indices[:9]
    n, t = i // T, i % T
    np.testing.assert_array_equal(batch.observations.cpu().numpy(), buf.observations[t, n].cpu().numpy())
    np.testing.assert_array_equal(batch.advantages.cpu().numpy(), buf.advantages[t, n, 0].cpu().numpy())


def test_rollout_observations_are_born_in_the_buffer_slots():
    """with `bind_rollout_slots` the env writes step t's observation straight into buffer slot t+1 (add() then
    skips its copy); the buffer content is identical to the copying path."""
    g = EnvGolden("env_g20_long")
    T = 10
    bufs = []
    for bind in (True, False):
        env = EnvWrapperGenNBVTrain(make_env(g))
        algo, _ = make_algo(env, n_steps=T, batch_size=9, n_epochs=1)
        algo.bind_rollout_slots = bind
        algo._setup_learn()
        ptrs = []
        assert algo.collect_rollouts(callback=lambda loc: ptrs.append(loc["new_obs"].data_ptr()))
        buf = algo.rollout_buffer
        slots = [buf.observations[t].data_ptr() for t in range(T)]
        if bind:
            assert ptrs[:T - 1] == slots[1:], "observations of steps 0..T-2 must live in slots 1..T-1"
            assert ptrs[T - 1] not in slots               # the last one stays in the env's own buffer for the next rollout
        else:
            assert not set(ptrs) & set(slots)
        bufs.append({k: getattr(buf, k).clone() for k in ("observations", "actions", "rewards", "episode_starts", "values",
                                                          "log_probs", "advantages", "returns")})
    for k in bufs[0]:
        assert torch.equal(bufs[0][k], bufs[1][k]), k


@pytest.mark.parametrize("target_kl", [None, 1e-7])
def test_fused_train_matches_torch_rerun(target_kl):
    g = EnvGolden("env_g20_long")
    env = EnvWrapperGenNBVTrain(make_env(g))
    T, B, E = 8, 12, 2
    algo, ref = make_algo(env, n_steps=T, batch_size=B, n_epochs=E, target_kl=target_kl)
    algo._setup_learn()
    algo.collect_rollouts()
    buf = algo.rollout_buffer
    # ---- plain torch re-run of ppo_grid_obs.py:176-297 on the CPU with the same buffer content and permutation
    N = g.N
    flat = lambda x: x.transpose(0, 1).reshape(N * T, *x.shape[2:]).cpu()          # swap_and_flatten (buffers.py:56-69)
    obs, acts = flat(buf.observations), flat(buf.actions).long()
    vals, lps, advs, rets = (flat(x).flatten() for x in (buf.values, buf.log_probs, buf.advantages, buf.returns))
    opt = torch.optim.Adam(ref.parameters(), lr=1e-4, eps=1e-5)
    ref.train()
    logs, stop = [], False
    for epoch in range(E):
        for start in range(0, N * T, B):
            idx = buf.indices[start:start + B]
            v, lp, ent = ref.evaluate_actions(obs[idx], acts[idx])
            loss, parts = encoder_ref.ppo_loss(v, lp, ent, vals[idx], lps[idx], advs[idx], rets[idx])
            logs.append([float(loss.detach())] + [float(parts[k].detach()) for k in
                                                  ("policy_loss", "value_loss", "entropy_loss", "approx_kl", "clip_fraction")])
            if target_kl is not None and float(parts["approx_kl"]) > 1.5 * target_kl:
                stop = True
                break
            opt.zero_grad()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
            opt.step()
        if stop:
            break
    before = {k: v.clone() for k, v in algo.policy.state_dict().items()}
    algo.train()
    logs = np.array(logs)
    rec = algo.logger.name_to_value
    for key, col in (("train/policy_gradient_loss", 1), ("train/value_loss", 2), ("train/entropy_loss", 3),
                     ("train/approx_kl", 4), ("train/clip_fraction", 5)):
        assert abs(rec[key] - logs[:, col].mean()) <= 2e-4 * max(1.0, abs(logs[:, col].mean())), key
    assert abs(rec["train/loss"] - logs[-1, 0]) <= 2e-4 * max(1.0, abs(logs[-1, 0]))
    steps = len(logs) - (1 if stop else 0)
    assert algo._adam_step == steps
    sd_ref = ref.state_dict()
    for k, v in algo.policy.state_dict().items():
        a, b = v.detach().cpu().double(), sd_ref[k].double()
        if k.endswith("num_batches_tracked"):
            assert int(a) == int(b), k
            continue
        if k.endswith("naive_encoder_grid.0.bias") or k.endswith("naive_encoder_grid.3.bias"):
            continue          # zero-gradient parameters (conv bias under batch-stat BN): Adam amplifies rounding noise
        # parameters moved by ~lr per step: compare the displacement, not the O(1) values
        disp = (b - before[k].cpu().double()).abs().max()
        assert float((a - b).abs().max()) <= 0.02 * float(disp) + 1e-7, (k, float((a - b).abs().max()), float(disp))

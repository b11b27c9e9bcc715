"""TEST HARNESS ONLY (build container: needs /root/reference) -- drives the reference's OWN
`PPO_Grid_Obs.collect_rollouts()` + `train()` (stable_baselines3/ppo/ppo_grid_obs.py:176-297,
stable_baselines3/common/on_policy_algorithm_grid_obs.py:128-221) and its
`TensorRolloutBuffer_Grid_Obs.compute_returns_and_advantage` (stable_baselines3/common/buffers.py:706-724)
on the CPU with a scripted environment, so that the restatements used as oracles on the GPU box
(oracle/gennbv_oracle.c::gae, oracle/encoder_ref.py::PolicyRef + ppo_loss + torch Adam) are pinned
against the reference itself, and writes tests/golden/ppo_train_g20.npz.

    python oracle/ref_ppo_driver.py            # (re)writes tests/golden/ppo_train_g20.npz

Nothing here is reachable from the product package.
"""
import os
import sys
from collections import deque

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import encoder_ref  # noqa: E402
import ref_loader  # noqa: E402

G, STATE, RGB = 20, 600, 8192
D = STATE + G ** 3 + RGB
NVEC = (81, 81, 51, 1, 13, 13)
PPO_KW = dict(learning_rate=1e-4, gamma=0.99, gae_lambda=0.95, clip_range=0.2, clip_range_vf=0.2, ent_coef=0.01,
              vf_coef=0.8, max_grad_norm=1)          # train_gennbv.py:170-187
SAMPLE_STRIDE, FULL_LIMIT = 37, 70000


def scripted_observations(n_envs, steps, seed):
    """[steps+1, N, D] observations (pose history, tri-class grid, two gray frames), rewards, dones, time-outs."""
    g = torch.Generator().manual_seed(seed)
    obs = torch.cat([torch.randn(steps + 1, n_envs, STATE, generator=g),
                     torch.randint(-1, 2, (steps + 1, n_envs, G ** 3), generator=g).float(),
                     torch.randint(0, 4, (steps + 1, n_envs, RGB), generator=g).float() / 4], dim=2)
    rew = torch.randn(steps, n_envs, generator=g)
    done = torch.rand(steps, n_envs, generator=g) < 0.2
    tout = done & (torch.rand(steps, n_envs, generator=g) < 0.5)
    return obs, rew, done, tout


def make_scripted_env(ref, n_envs, steps, seed):
    """A scripted env behind the reference's own wrapper class (so that `is_isaac_gym_env` is True,
    stable_baselines3/utils.py:23-35)."""
    from gym import spaces
    obs, rew, done, tout = scripted_observations(n_envs, steps, seed)

    class Scripted(ref.wrapper.EnvWrapperGenNBVTrain):
        def __init__(self):
            self.observation_space = spaces.Box(low=-np.inf, high=np.inf, shape=(D,), dtype=np.float32)
            self.action_space = spaces.MultiDiscrete(list(NVEC))
            self.num_envs, self.t = n_envs, 0
            self.max_episode_length = 30
            self.episode_length_buf = torch.zeros(n_envs, dtype=torch.long)
            self._gym_env = None

        def seed(self, s):
            pass

        def reset(self):
            self.t = 0
            return obs[0].clone()

        def step(self, actions):
            t = self.t
            self.t += 1
            return obs[t + 1].clone(), rew[t].clone(), done[t].clone(), {"time_outs": tout[t].clone(), "episode": {}}

    return Scripted()


class _Log:
    def __init__(self):
        self.name_to_value = {}

    def record(self, k, v, exclude=None):
        self.name_to_value[k] = v

    def dump(self, step=0):
        pass


def make_reference_algo(ref, env, n_steps, batch_size, n_epochs, target_kl, weight_seed, torch_seed):
    policy_kwargs = dict(net_arch=[], features_extractor_class=ref.encoder.Hybrid_Encoder, features_extractor_kwargs=dict(
        encoder_param={"hidden_shapes": [256, 256], "visual_dim": 256},
        net_param={"transformer_params": [[1, 256], [1, 256]], "append_hidden_shapes": [256, 256]},
        state_input_shape=(STATE,), visual_input_shape=(100, 48, 48)))
    algo = ref.ppo.PPO_Grid_Obs(ref.policies.ActorCriticPolicy_Train_Eval, env, n_steps=n_steps, batch_size=batch_size,
                                n_epochs=n_epochs, target_kl=target_kl, policy_kwargs=policy_kwargs, device="cpu",
                                seed=torch_seed, **PPO_KW)
    mirror = encoder_ref.PolicyRef(G, STATE)
    sd = encoder_ref.seeded_state_dict(mirror, weight_seed, scale=0.5)
    algo.policy.load_state_dict(sd)
    mirror.load_state_dict(sd)
    algo._logger = _Log()
    algo.ep_info_buffer = deque(maxlen=100)
    algo._last_obs = env.reset()
    algo._last_episode_starts = np.ones((env.num_envs,), dtype=bool)
    return algo, mirror, sd


def run_reference(n_envs=4, n_steps=8, batch_size=8, n_epochs=2, target_kl=None, weight_seed=3, torch_seed=5, env_seed=9):
    """Runs the reference's collect_rollouts() + train(); returns everything a replay needs."""
    ref = ref_loader.load_reference()
    from stable_baselines3.common.callbacks import BaseCallback

    class Quiet(BaseCallback):
        def _on_step(self):
            return True

    env = make_scripted_env(ref, n_envs, n_steps, env_seed)
    algo, mirror, sd = make_reference_algo(ref, env, n_steps, batch_size, n_epochs, target_kl, weight_seed, torch_seed)
    cb = Quiet()
    cb.init_callback(algo)
    assert algo.collect_rollouts(env, cb, algo.rollout_buffer, n_rollout_steps=n_steps)
    buf = algo.rollout_buffer
    cols = {k: getattr(buf, k).clone() for k in ("observations", "actions", "rewards", "episode_starts", "values", "log_probs",
                                                  "advantages", "returns")}          # [T,N,...] (before swap_and_flatten)
    indices = buf.indices.copy()
    last_values = algo.policy.predict_values(algo._last_obs).detach().clone()
    last_dones = torch.as_tensor(algo._last_episode_starts).clone()
    algo.train()
    return dict(ref=ref, algo=algo, mirror=mirror, state_dict=sd, cols=cols, indices=indices, last_values=last_values,
                last_dones=last_dones, logs=dict(algo.logger.name_to_value),
                after={k: v.detach().clone() for k, v in algo.policy.state_dict().items()},
                cfg=dict(n_envs=n_envs, n_steps=n_steps, batch_size=batch_size, n_epochs=n_epochs, target_kl=target_kl,
                         weight_seed=weight_seed))


def mirror_train(mirror, cols, indices, cfg):
    """The restatement the GPU tests use as oracle: PolicyRef + ppo_loss + torch Adam on the same buffer and permutation."""
    T, N, B = cfg["n_steps"], cfg["n_envs"], cfg["batch_size"]
    flat = lambda x: x.transpose(0, 1).reshape(N * T, *x.shape[2:])                  # swap_and_flatten (buffers.py:56-69)
    obs, acts = flat(cols["observations"]), flat(cols["actions"]).long()
    vals, lps, advs, rets = (flat(cols[k]).flatten() for k in ("values", "log_probs", "advantages", "returns"))
    opt = torch.optim.Adam(mirror.parameters(), lr=PPO_KW["learning_rate"], eps=1e-5)
    mirror.train()
    logs, stop, steps = [], False, 0
    for epoch in range(cfg["n_epochs"]):
        kls = []
        for start in range(0, N * T, B):
            idx = indices[start:start + B]
            v, lp, ent = mirror.evaluate_actions(obs[idx], acts[idx])
            loss, parts = encoder_ref.ppo_loss(v, lp, ent, vals[idx], lps[idx], advs[idx], rets[idx])
            logs.append([float(loss.detach())] + [float(parts[k].detach()) for k in
                                                  ("policy_loss", "value_loss", "entropy_loss", "approx_kl", "clip_fraction")])
            kls.append(float(parts["approx_kl"]))
            if cfg["target_kl"] is not None and float(parts["approx_kl"]) > 1.5 * cfg["target_kl"]:
                stop = True
                break
            opt.zero_grad()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(mirror.parameters(), PPO_KW["max_grad_norm"])
            opt.step()
            steps += 1
        if stop:
            break
    logs = np.array(logs)
    return dict(logs=logs, last_epoch_kl=float(np.mean(kls)), steps=steps)


def compact_params(sd):
    """Small tensors whole, large ones as a strided sample + float64 sum (keeps the fixture under 1 MB)."""
    out = {}
    for k, v in sd.items():
        a = v.detach().cpu().numpy().reshape(-1)
        if a.size <= FULL_LIMIT:
            out["after/" + k] = a
        else:
            out["after_sample/" + k] = a[::SAMPLE_STRIDE].copy()
            out["after_sum/" + k] = np.array(a.astype(np.float64).sum())
    return out


def write_golden(path):
    r = run_reference()
    c = r["cols"]
    out = dict(meta=np.array([r["cfg"][k] for k in ("n_envs", "n_steps", "batch_size", "n_epochs", "weight_seed")]),
               state=c["observations"][..., :STATE].numpy(), grid=c["observations"][..., STATE:STATE + G ** 3].numpy().astype(np.int8),
               rgb4=(c["observations"][..., STATE + G ** 3:] * 4).numpy().astype(np.uint8), actions=c["actions"].numpy(), rewards=c["rewards"].numpy(),
               episode_starts=c["episode_starts"].numpy(), values=c["values"].numpy(), log_probs=c["log_probs"].numpy(),
               advantages=c["advantages"].numpy(), returns=c["returns"].numpy(), indices=r["indices"],
               last_values=r["last_values"].numpy(), last_dones=r["last_dones"].numpy().astype(np.uint8),
               log_keys=np.array(sorted(k for k in r["logs"] if k.startswith("train/"))),
               log_vals=np.array([float(r["logs"][k]) for k in sorted(k for k in r["logs"] if k.startswith("train/"))]))
    out.update(compact_params(r["after"]))
    np.savez_compressed(path, **out)
    return path


if __name__ == "__main__":
    p = write_golden(os.path.join(ROOT, "tests", "golden", "ppo_train_g20.npz"))
    print(p, os.path.getsize(p), "bytes")

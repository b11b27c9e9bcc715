"""TEST INFRASTRUCTURE (checker side): replays tests/golden/ppo_train_g20.npz -- a rollout buffer filled by the REFERENCE's
`collect_rollouts()`, its permutation, and what the reference's `train()` logged and left in the policy
(written by oracle/ref_ppo_driver.py in the build container) -- through gennbv_b200's fused PPO update on the GPU and compares
within the 1e-4 budget.  Used by tests/test_ppo_gpu.py and by __graft_entry__.smoke()."""
import os

import numpy as np
import torch

import encoder_ref

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ppo_train_g20.npz")


def replay_ppo_golden(device="cuda:0", use_cuda_graph=True):
    from gennbv_b200.ppo import PPO_Grid_Obs
    from gennbv_b200.spaces import Box, MultiDiscrete
    d = np.load(GOLDEN)
    N, T, B, E, wseed = (int(v) for v in d["meta"])
    G = 20
    obs = np.concatenate([d["state"], d["grid"].astype(np.float32), d["rgb4"].astype(np.float32) / 4], axis=2)

    class Stub:                      # the algorithm only needs the spaces and the env count to build its buffers
        observation_space = Box(-np.inf, np.inf, (obs.shape[2],), np.float32)
        action_space = MultiDiscrete([81, 81, 51, 1, 13, 13])
        num_envs, grid_size = N, G

        def seed(self, s):
            pass

    kw = dict(net_arch=[], features_extractor_kwargs=dict(
        encoder_param={"hidden_shapes": [256, 256], "visual_dim": 256},
        net_param={"transformer_params": [[1, 256], [1, 256]], "append_hidden_shapes": [256, 256]},
        state_input_shape=(600,), visual_input_shape=(100, 48, 48)))
    algo = PPO_Grid_Obs(env=Stub(), learning_rate=1e-4, n_steps=T, batch_size=B, n_epochs=E, gamma=0.99, gae_lambda=0.95,
                        clip_range=0.2, clip_range_vf=0.2, ent_coef=0.01, vf_coef=0.8, max_grad_norm=1, target_kl=None,
                        policy_kwargs=kw, seed=0, device=device)
    algo.use_cuda_graph = use_cuda_graph
    ref = encoder_ref.PolicyRef(G, 600)
    algo.policy.load_state_dict(encoder_ref.seeded_state_dict(ref, wseed, scale=0.5))
    buf = algo.rollout_buffer
    buf.reset()
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    buf.observations.copy_(to(obs))
    for k in ("actions", "rewards", "values", "log_probs", "advantages", "returns"):
        getattr(buf, k).copy_(to(d[k]))
    buf.episode_starts.copy_(to(d["episode_starts"]))
    buf.set_permutation(d["indices"])
    buf.step = buf.pos = T
    buf.full = True
    # GAE of the reference's stored columns is the kernel's, bit for bit
    adv, ret = buf.advantages.clone(), buf.returns.clone()
    buf.compute_returns_and_advantage(last_values=to(d["last_values"]), dones=to(d["last_dones"]).bool())
    assert torch.equal(buf.advantages, adv) and torch.equal(buf.returns, ret), "GAE differs from the reference's buffer"
    algo.train()
    rec = algo.logger.name_to_value
    want = dict(zip([str(k) for k in d["log_keys"]], d["log_vals"]))
    worst = 0.0
    for k in ("train/entropy_loss", "train/policy_gradient_loss", "train/value_loss", "train/approx_kl", "train/clip_fraction",
              "train/loss", "train/explained_variance"):
        err = abs(rec[k] - want[k]) / max(1.0, abs(want[k]))
        worst = max(worst, err)
        assert err <= 1e-4, (k, rec[k], want[k])
    assert rec["train/n_updates"] == want["train/n_updates"] and rec["train/clip_range"] == want["train/clip_range"]
    assert algo._adam_step == E * (-(-N * T // B))
    for k, v in algo.policy.state_dict().items():
        a = v.detach().cpu().numpy().reshape(-1).astype(np.float64)
        if "after/" + k in d.files:
            b = d["after/" + k].astype(np.float64)
        else:
            a, b = a[::37], d["after_sample/" + k].astype(np.float64)
        err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)
        worst = max(worst, float(err))
        assert err <= 1e-4, (k, err)
    return dict(worst_rel_err=worst, optimizer_steps=algo._adam_step, logs={k: rec[k] for k in want if k in rec})

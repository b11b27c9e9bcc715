"""Worker of tests/test_ppo_multi_gpu.py: one rank of a 2-GPU NCCL run of `PPO_Grid_Obs.train()` on the golden rollout buffer
(tests/golden/ppo_train_g20.npz).  Launched with torch.distributed.run; rank 0 prints one JSON line.

  scenario "same":  both ranks hold the same buffer -> the averaged gradient equals each rank's own, so the 2-rank run must
                    reproduce the single-GPU run bit for bit (any race in the overlapped all-reduce would show);
  scenario "vote":  rank 1's stored log-probs are perturbed so that only ITS approx_kl exceeds 1.5 * target_kl at the first
                    minibatch -> both ranks must stop there (no Adam step, no hang), via the vote carried by the all-reduce.
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def digest(t):
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()


def build(dev, target_kl, perturb):
    import encoder_ref
    from gennbv_b200.ppo import PPO_Grid_Obs
    from gennbv_b200.spaces import Box, MultiDiscrete
    d = np.load(os.path.join(HERE, "golden", "ppo_train_g20.npz"))
    N, T, B, E, wseed = (int(v) for v in d["meta"])
    obs = np.concatenate([d["state"], d["grid"].astype(np.float32), d["rgb4"].astype(np.float32) / 4], axis=2)

    class Stub:
        observation_space = Box(-np.inf, np.inf, (obs.shape[2],), np.float32)
        action_space = MultiDiscrete([81, 81, 51, 1, 13, 13])
        num_envs, grid_size = N, 20

        def seed(self, s):
            pass

    kw = dict(net_arch=[], features_extractor_kwargs=dict(
        encoder_param={"hidden_shapes": [256, 256], "visual_dim": 256},
        net_param={"transformer_params": [[1, 256], [1, 256]], "append_hidden_shapes": [256, 256]},
        state_input_shape=(600,), visual_input_shape=(100, 48, 48)))
    algo = PPO_Grid_Obs(env=Stub(), learning_rate=1e-4, n_steps=T, batch_size=B, n_epochs=E, gamma=0.99, gae_lambda=0.95,
                        clip_range=0.2, clip_range_vf=0.2, ent_coef=0.01, vf_coef=0.8, max_grad_norm=1, target_kl=target_kl,
                        policy_kwargs=kw, seed=0, device=dev)
    algo.policy.load_state_dict(encoder_ref.seeded_state_dict(encoder_ref.PolicyRef(20, 600), wseed, scale=0.5))
    buf = algo.rollout_buffer
    buf.reset()
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    buf.observations.copy_(to(obs))
    for k in ("actions", "rewards", "values", "log_probs", "advantages", "returns"):
        getattr(buf, k).copy_(to(d[k]))
    if perturb:
        buf.log_probs.add_(0.7)
    buf.episode_starts.copy_(to(d["episode_starts"]))
    buf.set_permutation(d["indices"])
    buf.step = buf.pos = T
    buf.full = True
    return algo


def summary(algo):
    pol = algo.policy
    return {"params": digest(pol.flat_params), "m": digest(algo._exp_avg), "v": digest(algo._exp_avg_sq),
            "bn": [digest(b) for b in pol.buffers()], "adam_step": algo._adam_step,
            "logged": algo._last_train["minibatches_logged"], "stopped_epoch": algo._last_train["stopped_epoch"]}


def main():
    scenario = sys.argv[1]
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank, local = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    if scenario == "same":
        algo = build(dev, None, perturb=False)
    else:
        # threshold between the two ranks' first-minibatch approx_kl: rank 0 ~1e-3..1e-2 (train- vs eval-mode BN), rank 1 ~0.2
        algo = build(dev, 0.05, perturb=(rank == 1))
    algo.train()
    torch.cuda.synchronize()
    mine = summary(algo)
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        dist.destroy_process_group()
    else:
        gathered = [mine]
    if rank == 0:
        print("RESULT " + json.dumps(gathered))


if __name__ == "__main__":
    main()

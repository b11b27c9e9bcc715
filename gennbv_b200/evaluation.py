"""`evaluate_policy_grid_obs` / `AUC_update` -- the evaluation loop of the reference
(stable_baselines3/common/evaluation.py:136-378) for the eval env's five-element interface.

Same results, B200-first mechanics: the running rewards / lengths / AUC table live on the device and are updated for
all envs at once (the reference walks a Python loop over the 50 envs every step, :288-329 and :370-376, with a host
read per env); one host read of the done flags per step decides the bookkeeping.

The reference hard-codes `n_envs = 50` and `max_length = 30` "to align with args_eval.num_envs" (:199-202); here they
are read from the env (`num_envs`, `max_episode_length`), which gives the same numbers for the reference's eval setup.
"""
import numpy as np
import torch


def AUC_update(AUC_rews, cur_rewards, cur_lengths, dones, episode_done_flag):
    """evaluation.py:356-378 for all envs at once.
    AUC_rews [n_envs, max_length]; cur_rewards [n_envs]; cur_lengths: the global step count (1-based scalar);
    dones [n_envs]; episode_done_flag [n_envs] (non-zero once the env's episode has ended on an earlier step)."""
    col = int(cur_lengths) - 1
    finished = episode_done_flag.to(AUC_rews.device) != 0
    running = dones.to(AUC_rews.device) == 0
    prev = AUC_rews[:, col - 1]                       # col - 1 == -1 wraps to the last column exactly like the reference's index
    AUC_rews[:, col] = torch.where(finished, prev, torch.where(running, cur_rewards.to(AUC_rews), AUC_rews[:, col]))
    return AUC_rews


def evaluate_policy_grid_obs(model, env, n_eval_episodes=10, deterministic=True, render=False, callback=None,
                             reward_threshold=None, return_episode_rewards=False, return_AUC=True, return_Accuracy=True,
                             warn=True):
    """Runs `model.predict` on `env` (an `EnvWrapperGenNBVEval`) until every env has finished its share of
    `n_eval_episodes` episodes.  Returns, like the reference (:340-347):
      return_Accuracy : (episode_rewards, episode_lengths, mean_AUC [n_envs], episode_accuracies)
      return_AUC      : (episode_rewards, episode_lengths, mean_AUC)
      return_episode_rewards : (episode_rewards, episode_lengths)
      otherwise       : (mean_reward, std_reward)
    episode_rewards / episode_lengths are lists of 0-dim CPU tensors in the order episodes finish (env order within a
    step), episode_accuracies a list of floats."""
    n_envs = int(env.num_envs)
    max_length = int(env.max_episode_length)
    dev = env.device
    episode_rewards, episode_lengths, episode_accuracies = [], [], []
    targets = torch.tensor([(n_eval_episodes + i) // n_envs for i in range(n_envs)], dtype=torch.int64)
    counts = torch.zeros(n_envs, dtype=torch.int64)
    current_rewards = torch.zeros(n_envs, device=dev)
    current_lengths = torch.zeros(n_envs, dtype=torch.int64, device=dev)
    observations, rewards, dones, infos, accuracies = env.reset()
    if return_AUC:
        assert int(targets.max()) <= 1                                  # :267
        AUC_rews = torch.zeros(n_envs, max_length, device=dev)
        episode_done_flag = torch.zeros(n_envs, device=dev)
    global_length = 0
    while bool((counts < targets).any()):
        global_length += 1
        actions, _ = model.predict(observations, state=None, deterministic=deterministic)
        observations, rewards, dones, infos, accuracies = env.step(actions)
        if return_AUC:
            AUC_rews = AUC_update(AUC_rews, rewards, global_length, dones, episode_done_flag)
        current_rewards += rewards
        current_lengths += 1
        active = (counts < targets).to(dev)
        if return_AUC:
            episode_done_flag += (dones.to(dev) & active).float()      # only envs still being counted (:276-283)
        if callback is not None:
            callback(locals(), globals())
        finished = (dones.to(dev) & active)
        ids = finished.nonzero().flatten().tolist()                     # one host read per step
        if ids:
            r_host, l_host = current_rewards[ids].cpu(), current_lengths[ids].cpu()
            for k, i in enumerate(ids):
                if return_AUC:
                    episode_accuracies.append(accuracies[str(i)])
                episode_rewards.append(r_host[k].clone())
                episode_lengths.append(l_host[k].clone())
                counts[i] += 1
            idx = torch.tensor(ids, device=dev)
            current_rewards[idx] = 0
            current_lengths[idx] = 0
        if render:
            env.render()
    mean_AUC = None
    if return_AUC:
        weights = (max_length - torch.arange(max_length, device=dev)).to(AUC_rews)
        mean_AUC = (AUC_rews * weights).sum(dim=1) / max_length         # :335
    mean_reward = float(np.mean([float(r) for r in episode_rewards])) if episode_rewards else float("nan")
    std_reward = float(np.std([float(r) for r in episode_rewards])) if episode_rewards else float("nan")
    if reward_threshold is not None:
        assert mean_reward > reward_threshold, f"Mean reward below threshold: {mean_reward:.2f} < {reward_threshold:.2f}"
    if return_Accuracy:
        return episode_rewards, episode_lengths, mean_AUC, episode_accuracies
    if return_AUC:
        return episode_rewards, episode_lengths, mean_AUC
    if return_episode_rewards:
        return episode_rewards, episode_lengths
    return mean_reward, std_reward

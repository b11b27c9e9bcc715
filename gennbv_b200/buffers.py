"""`TensorRolloutBuffer_Grid_Obs` -- device-resident rollout buffer (stable_baselines3/common/buffers.py:628-762).

Same interface: `reset()`, `add(obs, action, reward, episode_start, value, log_prob)`,
`compute_returns_and_advantage(last_values, dones)`, `get(batch_size)` yielding `RolloutBufferSamples`, attributes
`observations / actions / rewards / episode_starts / values / log_probs / advantages / returns / buffer_size / n_envs`.

B200-first differences:
  * storage is allocated once ([T,N,D] fp32 = 35.5 GB at 64^3 fits the 180 GB HBM3e) instead of being re-created by
    every reset() (buffers.py:659-673);
  * `add()` skips the observation copy when the env already wrote the observation into this slot
    (`Env_Train_GenNBV.bind_next_observation`, used by `PPO_Grid_Obs.collect_rollouts`);
  * GAE is one kernel launch (gnbv_gae) instead of a Python loop of T x 8 launches (:706-724);
  * `get()` does not materialise swap_and_flatten's transposed copy of the observations (:56-69,736-746): the
    reference's env-major flat index i = n*T + t is mapped to the storage row t*N + n, and `minibatch_rows()` hands
    those rows to the encoder kernels, which read the minibatch in place (no 139 MB gather per minibatch, :753-762).
    `get()` itself still returns gathered tensors for callers that want the reference's sample tuples.
"""
from typing import NamedTuple

import numpy as np
import torch

from . import ops


class RolloutBufferSamples(NamedTuple):
    observations: torch.Tensor
    actions: torch.Tensor
    old_values: torch.Tensor
    old_log_prob: torch.Tensor
    advantages: torch.Tensor
    returns: torch.Tensor


class TensorRolloutBuffer_Grid_Obs:
    def __init__(self, buffer_size, observation_space, action_space, device="cuda", gae_lambda=1, gamma=0.99, n_envs=1):
        self.buffer_size, self.n_envs = int(buffer_size), int(n_envs)
        self.num_transitions_per_env, self.num_envs = self.buffer_size, self.n_envs
        self.obs_shape = tuple(observation_space.shape)
        self.actions_shape = int(action_space.shape[0])
        self.device = torch.device(device)
        # storage may live on any device (construction-only wiring tests); GAE and the minibatch kernels refuse non-CUDA tensors
        self.gae_lambda, self.gamma = gae_lambda, gamma
        T, N, dev = self.buffer_size, self.n_envs, self.device
        self.observations = torch.zeros(T, N, *self.obs_shape, device=dev)
        self.actions = torch.zeros(T, N, self.actions_shape, device=dev)
        self.rewards = torch.zeros(T, N, 1, device=dev)
        self.episode_starts = torch.zeros(T, N, 1, dtype=torch.uint8, device=dev)
        self.log_probs = torch.zeros(T, N, 1, device=dev)
        self.values = torch.zeros(T, N, 1, device=dev)
        self.returns = torch.zeros(T, N, 1, device=dev)
        self.advantages = torch.zeros(T, N, 1, device=dev)
        self.storage_rows = torch.zeros(T * N, dtype=torch.int64, device=dev)     # fixed address: captured CUDA graphs read it
        self.reset()

    def reset(self):
        self.step = self.pos = 0
        self.full = self.generator_ready = False
        self.set_permutation(np.random.permutation(self.buffer_size * self.n_envs))    # buffers.py:673, one per rollout

    def set_permutation(self, indices):
        """`indices`: a permutation of the reference's env-major flat index i = n*T + t (swap_and_flatten order); the
        storage row of i is t*N + n."""
        self.indices = np.asarray(indices)
        i = torch.from_numpy(self.indices).long()
        self.storage_rows.copy_((i % self.buffer_size) * self.n_envs + i // self.buffer_size)

    def add(self, obs, action, reward, episode_start, value, log_prob):
        if isinstance(episode_start, np.ndarray):
            episode_start = torch.from_numpy(episode_start).to(self.device)
        if self.step >= self.buffer_size:
            raise AssertionError("Rollout buffer overflow")
        t = self.step
        if obs.data_ptr() != self.observations[t].data_ptr():      # already in place when the env wrote it there (8f-2)
            self.observations[t].copy_(obs)
        self.actions[t].copy_(action)
        self.rewards[t].copy_(reward.view(-1, 1))
        self.episode_starts[t].copy_(episode_start.view(-1, 1))
        self.values[t].copy_(value)
        self.log_probs[t].copy_(log_prob.view(-1, 1))
        self.step += 1
        self.pos += 1
        self.full = self.pos == self.buffer_size

    def compute_returns_and_advantage(self, last_values, dones):
        T, N = self.buffer_size, self.n_envs
        ops.gae(self.rewards.view(T, N), self.values.view(T, N), self.episode_starts.view(T, N),
                last_values.detach().reshape(N).contiguous().float(), dones.reshape(N).to(torch.uint8).contiguous(),
                self.gamma, self.gae_lambda, self.advantages.view(T, N), self.returns.view(T, N))

    # ---- minibatches -----------------------------------------------------------------------------------------------
    def flat(self, name):
        """[T*N, ...] view of a stored tensor in storage order (row = t*N + n)."""
        x = getattr(self, name)
        return x.view(self.buffer_size * self.n_envs, *x.shape[2:])

    def minibatch_rows(self, batch_size):
        """Yields device index tensors (storage rows) of successive minibatches, same order as the reference's get()."""
        assert self.step == self.buffer_size, ""
        total = self.buffer_size * self.n_envs
        for start in range(0, total, batch_size):
            yield self.storage_rows[start:start + batch_size]

    def get(self, batch_size=None):
        assert self.step == self.buffer_size, ""
        if batch_size is None:
            batch_size = self.buffer_size * self.n_envs
        for rows in self.minibatch_rows(batch_size):
            yield RolloutBufferSamples(self.flat("observations")[rows], self.flat("actions")[rows],
                                       self.flat("values")[rows].flatten(), self.flat("log_probs")[rows].flatten(),
                                       self.flat("advantages")[rows].flatten(), self.flat("returns")[rows].flatten())

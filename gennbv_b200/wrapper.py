"""`EnvWrapperGenNBVTrain` / `EnvWrapperGenNBVEval` -- flatten the observation dict to [N, D] (gennbv/wrapper/env_wrapper_gennbv_train.py).

The reference concatenates `state | grid | state_rgb` into a new tensor every step (:27-56, one full copy of the
observation).  `gennbv_b200.Env_Train_GenNBV` already keeps its observation in that layout, so the wrapper returns
the env-owned `obs_flat` (zero copies); for any other dict-observation env it falls back to the same concatenation."""
import numpy as np
import torch

from .spaces import Box

KEY_SEQUENCE = ["state", "grid", "state_rgb"]


def flatten_observations(observation_dict, key_sequence):
    return torch.concat([observation_dict[k].reshape(observation_dict[k].shape[0], -1) for k in key_sequence], dim=-1)


def flatten_observation_spaces(observation_spaces, key_sequence):
    low = np.concatenate([np.asarray(observation_spaces.spaces[k].low).flatten() for k in key_sequence])
    high = np.concatenate([np.asarray(observation_spaces.spaces[k].high).flatten() for k in key_sequence])
    return Box(np.array(low, dtype=np.float32), np.array(high, dtype=np.float32), dtype=np.float32)


class EnvWrapperGenNBVTrain:
    def __init__(self, gym_env, observation_excluded=()):
        self.observation_excluded = observation_excluded
        self._gym_env = gym_env
        self.observation_space = flatten_observation_spaces(gym_env.observation_space, KEY_SEQUENCE)
        self.action_space = gym_env.action_space

    def __getattr__(self, attr):           # attribute reads fall through; assignments stay on the wrapper
        return getattr(self._gym_env, attr)

    def _flatten_observation(self, obs):
        flat = getattr(self._gym_env, "obs_flat", None)
        return flat if flat is not None else flatten_observations(obs, KEY_SEQUENCE)

    def reset(self):
        return self._flatten_observation(self._gym_env.reset())

    def step(self, action):
        obs, reward, done, info = self._gym_env.step(action)
        return self._flatten_observation(obs), reward, done, info

    def render(self, mode="human"):
        return self._gym_env.render(mode)

    def close(self):
        self._gym_env.close()


class EnvWrapperGenNBVEval(EnvWrapperGenNBVTrain):
    """gennbv/wrapper/env_wrapper_gennbv_eval.py:87-140: same flattening, five-element returns (accuracy dict last)."""

    def reset(self):
        obs, rews, dones, infos, accs = self._gym_env.reset()
        return self._flatten_observation(obs), rews, dones, infos, accs

    def step(self, action):
        obs, reward, done, info, accs = self._gym_env.step(action)
        return self._flatten_observation(obs), reward, done, info, accs

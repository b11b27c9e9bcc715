"""Minimal stand-in for the `gym` package (TEST HARNESS ONLY).

The reference (zjwzcx/GenNBV) imports `gym` (0.21-era API) in almost every
module; it is not installed in this image and there is no network.  This shim
provides just enough of the class surface for the reference's *own* Python to
import and run on CPU so that golden vectors can be generated from it
(oracle/gen_golden.py).  It is never imported by the product package.
"""
from . import spaces, utils, envs, wrappers, error, logger  # noqa: F401
from .spaces import Space  # noqa: F401

__version__ = "0.21.0"


class Env:
    metadata = {"render.modes": []}
    reward_range = (-float("inf"), float("inf"))
    spec = None
    action_space = None
    observation_space = None

    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def render(self, mode="human"):
        raise NotImplementedError

    def close(self):
        pass

    def seed(self, seed=None):
        return [seed]

    @property
    def unwrapped(self):
        return self


class GoalEnv(Env):
    pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self.action_space = env.action_space
        self.observation_space = env.observation_space
        self.reward_range = getattr(env, "reward_range", None)
        self.metadata = getattr(env, "metadata", {})

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def render(self, mode="human", **kwargs):
        return self.env.render(mode, **kwargs)

    def close(self):
        return self.env.close()

    def seed(self, seed=None):
        return self.env.seed(seed)

    @property
    def unwrapped(self):
        return self.env.unwrapped


class ObservationWrapper(Wrapper):
    def reset(self, **kwargs):
        return self.observation(self.env.reset(**kwargs))

    def step(self, action):
        obs, rew, done, info = self.env.step(action)
        return self.observation(obs), rew, done, info

    def observation(self, observation):
        raise NotImplementedError


class RewardWrapper(Wrapper):
    pass


class ActionWrapper(Wrapper):
    pass


def make(id, **kwargs):
    raise error.Error("gym shim: no registry")

"""TEST HARNESS ONLY -- imports the *unmodified* reference (zjwzcx/GenNBV) on CPU.

Used in the build container (where /root/reference exists) by oracle/gen_golden.py
and by the `-m "not gpu"` tests that pin the oracle restatement against the
reference's own Python.  Nothing here is reachable from the product package
`gennbv_b200`; nothing here runs on the GPU box (no /root/reference there).

How it works:
  * `oracle/ref_shim/gym` is a ~200-line stand-in for the absent `gym` package;
  * isaacgym / pycuda / open3d / pytorch3d / matplotlib / rsl_rl / wandb are
    replaced by MagicMock modules (none of them is touched by the functions the
    oracle drives: the PyCUDA Bresenham kernel cannot run on CPU and is restated
    in oracle/gennbv_oracle.c instead);
  * /root/reference is put on sys.path so `gennbv`, `stable_baselines3`,
    `legged_gym` resolve to the reference's own sources.
"""
import os
import sys
import types
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("GENNBV_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shim")

_MOCKED = [
    "open3d",
    "pycuda", "pycuda.driver", "pycuda.autoinit", "pycuda.compiler",
    "isaacgym", "isaacgym.gymapi", "isaacgym.gymutil", "isaacgym.gymtorch", "isaacgym.torch_utils",
    "matplotlib", "matplotlib.pyplot", "matplotlib.figure",
    "pytorch3d", "pytorch3d.loss",
    "rsl_rl", "rsl_rl.env", "rsl_rl.runners",
    "wandb", "wandb.sdk", "wandb.sdk.lib", "wandb.sdk.lib.telemetry",
]


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "gennbv"))


def load_reference():
    """Make `import gennbv`, `import stable_baselines3` resolve to the reference. Idempotent."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(1, REFERENCE_ROOT)
    for name in _MOCKED:
        if name not in sys.modules:
            m = MagicMock(name=name)
            m.__path__ = []  # behave like a package for "import a.b"
            m.__spec__ = None
            sys.modules[name] = m
    # `from isaacgym import *` / `from isaacgym.torch_utils import *` need __all__
    sys.modules["isaacgym"].__all__ = []
    # the real isaacgym.torch_utils star-exports numpy/torch; env_train_base.py relies on that `np`
    import numpy
    import torch
    tu = sys.modules["isaacgym.torch_utils"]
    tu.np, tu.torch = numpy, torch
    tu.__all__ = ["np", "torch"]
    # class bodies like `class X(VecEnv)` from rsl_rl need a real type
    sys.modules["rsl_rl.env"].VecEnv = type("VecEnv", (), {})
    import gennbv  # noqa: F401  (registers tasks; pulls in utils/env/network)
    return types.SimpleNamespace(
        gennbv=sys.modules["gennbv"],
        utils=__import__("gennbv.utils", fromlist=["x"]),
        env_train=__import__("gennbv.env.env_train_gennbv", fromlist=["x"]),
        env_base=__import__("gennbv.env.env_train_base", fromlist=["x"]),
        encoder=__import__("gennbv.network.hybrid_encoder", fromlist=["x"]),
        wrapper=__import__("gennbv.wrapper.env_wrapper_gennbv_train", fromlist=["x"]),
        buffers=__import__("stable_baselines3.common.buffers", fromlist=["x"]),
        policies=__import__("stable_baselines3.common.policies", fromlist=["x"]),
        ppo=__import__("stable_baselines3.ppo.ppo_grid_obs", fromlist=["x"]),
    )

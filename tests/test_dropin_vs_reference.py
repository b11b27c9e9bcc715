"""CPU, build container only (needs /root/reference): the drop-in claim, exercised.

1. Every mirrored class / function takes the reference's parameters: same names, same order, same defaults
   (`inspect.signature`), and exposes the reference's public methods.
2. The reference's own entry script, gennbv/train/train_gennbv.py, is imported UNMODIFIED with the `sys.modules` aliases of
   INTEGRATION.md section 1 installed and its `main()` is run up to `model.learn(...)`: argument parsing, `task_registry.make_env`
   (which constructs the env with the reference's five keyword arguments), the reference's wrapper call, the config dict,
   `ReconstructionCallBack` / `CallbackList`, `PPO_Grid_Obs(**config["algo"])`.  Construction only -- no kernel runs on this
   CPU box (and none may: there is no CPU compute path); `learn` is intercepted and its arguments are checked.
3. `PPO_Grid_Obs.save()` writes a zip the reference's `load_from_zip_file` reads, and `set_parameters()` reads a zip the
   reference's `save_to_zip_file` wrote.
"""
import inspect
import io
import os
import sys
import types
import zipfile

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.reference


def _ref():
    import ref_loader
    return ref_loader.load_reference()


def _params(fn):
    return [(p.name, p.default) for p in inspect.signature(fn).parameters.values()
            if p.kind in (p.POSITIONAL_OR_KEYWORD, p.KEYWORD_ONLY)]


def _assert_accepts(mine, theirs, what, skip=()):
    """Every parameter of the reference callable exists here, in the same relative order, with the same default."""
    m, t = _params(mine), [p for p in _params(theirs) if p[0] not in skip]
    names = [n for n, _ in m]
    last = -1
    for name, default in t:
        assert name in names, f"{what}: parameter `{name}` of the reference is missing"
        i = names.index(name)
        assert i > last, f"{what}: parameter `{name}` is out of order"
        last = i
        if default is not inspect.Parameter.empty and not callable(default):
            mine_default = m[i][1]
            assert mine_default == default or (isinstance(default, str) and default in ("auto", "cpu")), \
                f"{what}: default of `{name}` is {mine_default!r}, reference {default!r}"


def test_signatures_and_public_methods_mirror_the_reference():
    ref = _ref()
    import gennbv_b200.buffers as B
    import gennbv_b200.network as N
    import gennbv_b200.policy as P
    import gennbv_b200.ppo as A
    import gennbv_b200.utils as U
    import gennbv_b200.wrapper as W
    import gennbv_b200.env as E
    R = ref.ppo.PPO_Grid_Obs
    _assert_accepts(A.PPO_Grid_Obs.__init__, R.__init__, "PPO_Grid_Obs.__init__")
    _assert_accepts(A.PPO_Grid_Obs.learn, R.learn, "PPO_Grid_Obs.learn")
    _assert_accepts(A.PPO_Grid_Obs.collect_rollouts, R.collect_rollouts, "PPO_Grid_Obs.collect_rollouts")
    _assert_accepts(A.PPO_Grid_Obs.set_parameters, R.set_parameters, "PPO_Grid_Obs.set_parameters")
    _assert_accepts(A.PPO_Grid_Obs.save, R.save, "PPO_Grid_Obs.save")
    _assert_accepts(A.PPO_Grid_Obs.predict, R.predict, "PPO_Grid_Obs.predict")
    for name in ("train", "learn", "collect_rollouts", "predict", "save", "load", "set_parameters", "get_parameters", "get_env"):
        assert callable(getattr(A.PPO_Grid_Obs, name)), name
    RB = ref.buffers.TensorRolloutBuffer_Grid_Obs
    for fn in ("__init__", "add", "compute_returns_and_advantage", "get"):
        _assert_accepts(getattr(B.TensorRolloutBuffer_Grid_Obs, fn), getattr(RB, fn), f"TensorRolloutBuffer_Grid_Obs.{fn}")
    RP = ref.policies.ActorCriticPolicy_Train_Eval
    _assert_accepts(P.ActorCriticPolicy_Train_Eval.__init__, RP.__init__, "ActorCriticPolicy_Train_Eval.__init__",
                    skip=("use_sde", "log_std_init", "full_std", "sde_net_arch", "use_expln", "squash_output", "normalize_images"))
    for fn in ("forward", "evaluate_actions", "predict_values", "predict", "extract_features", "set_training_mode"):
        _assert_accepts(getattr(P.ActorCriticPolicy_Train_Eval, fn), getattr(RP, fn), f"ActorCriticPolicy_Train_Eval.{fn}")
    _assert_accepts(N.Hybrid_Encoder.__init__, ref.encoder.Hybrid_Encoder.__init__, "Hybrid_Encoder.__init__")
    _assert_accepts(N.Hybrid_Encoder.forward, ref.encoder.Hybrid_Encoder.forward, "Hybrid_Encoder.forward")
    RW = ref.wrapper.EnvWrapperGenNBVTrain
    for fn in ("__init__", "reset", "step", "close"):
        _assert_accepts(getattr(W.EnvWrapperGenNBVTrain, fn), getattr(RW, fn), f"EnvWrapperGenNBVTrain.{fn}")
    for fn in ("scanned_pts_to_idx_3D", "pose_coord_to_idx_3D", "bresenham3D_pycuda", "grid_occupancy_tri_cls"):
        _assert_accepts(getattr(U, fn), getattr(ref.utils, fn), f"gennbv.utils.{fn}")
    RE = ref.env_train.Env_Train_GenNBV
    for fn in ("step", "reset", "post_physics_step"):
        _assert_accepts(getattr(E.Env_Train_GenNBV, fn), getattr(RE, fn), f"Env_Train_GenNBV.{fn}")
    # BaseTask.__init__(cfg, sim_params, physics_engine, sim_device, headless) is what task_registry.make_env calls
    assert [n for n, _ in _params(E.Env_Train_GenNBV.__init__)][1:6] == ["cfg", "sim_params", "physics_engine", "sim_device", "headless"]


def _install_aliases():
    """INTEGRATION.md section 1."""
    import gennbv_b200.buffers, gennbv_b200.env, gennbv_b200.network, gennbv_b200.policy, gennbv_b200.ppo  # noqa: E401
    import gennbv_b200.utils, gennbv_b200.wrapper  # noqa: E401
    saved = {}
    alias = {"gennbv.network.hybrid_encoder": gennbv_b200.network,
             "gennbv.wrapper.env_wrapper_gennbv_train": gennbv_b200.wrapper,
             "gennbv.env.env_train_gennbv": gennbv_b200.env,
             "stable_baselines3.ppo.ppo_grid_obs": gennbv_b200.ppo}
    pol = types.ModuleType("stable_baselines3.common.policies")
    pol.ActorCriticPolicy_Train_Eval = gennbv_b200.policy.ActorCriticPolicy_Train_Eval
    alias["stable_baselines3.common.policies"] = pol
    for k, v in alias.items():
        saved[k] = sys.modules.get(k)
        sys.modules[k] = v
    return saved


def test_reference_train_script_runs_unmodified_up_to_learn(tmp_path, monkeypatch):
    ref = _ref()
    from gennbv_b200 import synth
    from gennbv_b200.env import Env_Train_GenNBV
    from gennbv_b200.ppo import PPO_Grid_Obs
    from gennbv_b200.sensors import SyntheticHouseSensor
    saved = _install_aliases()
    sys.modules.pop("gennbv.train.train_gennbv", None)
    try:
        # the task registry must hand out THIS env class with the reference's own config classes
        from legged_gym.utils import task_registry
        from gennbv.env.config_gennbv_train import Config_GenNBV_Train, DroneCfgPPO
        task_registry.register("train_gennbv", Env_Train_GenNBV, Config_GenNBV_Train, DroneCfgPPO)
        scenes = synth.make_house_scenes(2, 20, seed=0)
        monkeypatch.setattr(Env_Train_GenNBV, "grid_gt_loader", staticmethod(lambda env: scenes.grid_gt))
        monkeypatch.setattr(Env_Train_GenNBV, "sensor_factory",
                            staticmethod(lambda env: SyntheticHouseSensor(scenes.params, 32, 32, device=env.device)))
        import wandb_utils
        monkeypatch.setattr(wandb_utils, "team_name", "t", raising=False)
        monkeypatch.setattr(wandb_utils, "project_name", "p", raising=False)
        script = __import__("gennbv.train.train_gennbv", fromlist=["main"])
        assert script.PPO_Grid_Obs is PPO_Grid_Obs and script.Hybrid_Encoder.__module__ == "gennbv_b200.network"

        def fake_get_args(extra):
            ns = types.SimpleNamespace(task=None, num_envs=2, seed=3, headless=True, sim_device="cpu", rl_device="cpu",
                                       physics_engine="physx", stop_wandb=True, exp_name="t", horovod=False, resume=False,
                                       experiment_name=None, run_name=None, load_run=None, checkpoint=None, max_iterations=None,
                                       use_gpu=False, use_gpu_pipeline=False, subscenes=0, num_threads=0, device="cpu",
                                       compute_device_id=0, sim_device_type="cpu", sim_device_id=0, graphics_device_id=0)
            for a in extra:
                setattr(ns, a["name"].lstrip("-"), a.get("default"))
            ns.n_steps, ns.batch_size, ns.n_epochs = 4, 4, 2          # small buffers: this is a wiring test
            ns.save_freq = 10
            return ns

        monkeypatch.setattr(script, "get_args", fake_get_args)
        import legged_gym.utils.task_registry  # noqa: F401  (the attribute of that name on the package is the registry OBJECT)
        monkeypatch.setattr(sys.modules["legged_gym.utils.task_registry"], "parse_sim_params", lambda args, cfg: {"parsed": True})
        monkeypatch.setattr(script, "OPEN_ROBOT_ROOT_DIR", str(tmp_path))
        reached = {}

        def fake_learn(self, **kw):
            reached["kw"], reached["model"] = kw, self
            return self

        monkeypatch.setattr(PPO_Grid_Obs, "learn", fake_learn)
        script.main()
        model, kw = reached["model"], reached["kw"]
        assert isinstance(model, PPO_Grid_Obs) and isinstance(model.env._gym_env, Env_Train_GenNBV)
        assert model.env._gym_env.sim_params == {"parsed": True} and model.env.num_envs == 2
        assert kw["total_timesteps"] == 2 * 4 * script_total_iters(kw, 2 * 4) and kw["reset_num_timesteps"] is True
        assert kw["callback"].__class__.__name__ == "CallbackList"
        # the reference's callback list drives this algorithm object
        from gennbv_b200.ppo import _Callback
        cb = _Callback(kw["callback"], model)
        assert cb.objs and cb.objs[0].model is model
        assert model.n_steps == 4 and model.batch_size == 4 and model.target_kl == 0.05 and model.max_grad_norm == 1
        assert model.policy.features_extractor.grid_size == 20 and model.policy.optimizer.defaults["eps"] == 1e-5
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        sys.modules.pop("gennbv.train.train_gennbv", None)
        from gennbv.env.config_gennbv_train import Config_GenNBV_Train, DroneCfgPPO
        from legged_gym.utils import task_registry
        task_registry.register("train_gennbv", ref.env_train.Env_Train_GenNBV, Config_GenNBV_Train, DroneCfgPPO)


def script_total_iters(kw, per_iter):
    return kw["total_timesteps"] // per_iter


def _tiny_algo():
    from gennbv_b200.ppo import PPO_Grid_Obs
    from gennbv_b200.spaces import Box, MultiDiscrete
    D = 600 + 8000 + 8192

    class Stub:
        observation_space = Box(-np.inf, np.inf, (D,), np.float32)
        action_space = MultiDiscrete([81, 81, 51, 1, 13, 13])
        num_envs = 2

        def seed(self, s):
            pass

    kw = dict(net_arch=[], features_extractor_kwargs=dict(
        encoder_param={"hidden_shapes": [256, 256], "visual_dim": 256},
        net_param={"transformer_params": [[1, 256], [1, 256]], "append_hidden_shapes": [256, 256]},
        state_input_shape=(600,), visual_input_shape=(100, 48, 48)))
    return PPO_Grid_Obs(env=Stub(), n_steps=2, batch_size=2, n_epochs=1, policy_kwargs=kw, seed=1, device="cpu")


def test_checkpoints_round_trip_with_the_reference_zip_format(tmp_path):
    _ref()
    from stable_baselines3.common.save_util import load_from_zip_file, save_to_zip_file
    algo = _tiny_algo()
    algo._exp_avg.uniform_(-1, 1)
    algo._exp_avg_sq.uniform_(0, 1)
    algo._ctl[2] = 7
    path = algo.save(str(tmp_path / "ckpt"))
    # (a) the reference's loader reads our zip: data dict + the two state dicts under the reference's names
    data, params, pytorch_variables = load_from_zip_file(path, device="cpu")
    assert data["n_steps"] == 2 and set(params) == {"policy", "policy.optimizer"}
    sd = algo.policy.state_dict()
    assert list(params["policy"].keys()) == list(sd.keys())
    for k in sd:
        assert torch.equal(params["policy"][k], sd[k].cpu()), k
    st = params["policy.optimizer"]["state"]
    plist = list(algo.policy.parameters())
    assert len(st) == len(plist) and all(float(st[i]["step"]) == 7 for i in st)
    where = dict(zip(map(id, algo.policy.arena_parameters()), algo.policy._arena))
    for i, p in enumerate(plist):
        o, n = where[id(p)]
        assert torch.equal(st[i]["exp_avg"].reshape(-1), algo._exp_avg[o:o + n].cpu())
    # (b) we read a zip written by the reference's writer
    ref_zip = str(tmp_path / "ref_ckpt.zip")
    save_to_zip_file(ref_zip, data={"n_steps": 2}, params={"policy": params["policy"], "policy.optimizer": params["policy.optimizer"]},
                     pytorch_variables=None)
    other = _tiny_algo()
    other.policy.load_state_dict({k: torch.zeros_like(v) for k, v in sd.items()})
    other.set_parameters(ref_zip)
    for k, v in other.policy.state_dict().items():
        assert torch.equal(v, sd[k]), k
    for o, n in algo.policy._arena:                          # (the arenas also hold alignment padding between tensors)
        assert torch.equal(other._exp_avg[o:o + n], algo._exp_avg[o:o + n])
        assert torch.equal(other._exp_avg_sq[o:o + n], algo._exp_avg_sq[o:o + n])
    assert other._adam_step == 7

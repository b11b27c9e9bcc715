"""`PPO_Grid_Obs` -- the GenNBV PPO variant (stable_baselines3/ppo/ppo_grid_obs.py +
stable_baselines3/common/on_policy_algorithm_grid_obs.py) on the kernels of libgennbv_b200.

`collect_rollouts()` keeps the reference's control flow (policy forward in eval mode, env.step, time-out bootstrap
with the extra predict_values pass, buffer.add, the SB3 callback protocol) and `train()` keeps its arithmetic and
logged scalars, but a minibatch update is ONE fixed launch sequence of the C ABI (`gnbv_ppo_minibatch_grads` +
`gnbv_ppo_minibatch_apply`, csrc/ppo_update.cu) with no autograd graph and no host decision inside it:

    gather rows -> encoder forward (batch-stat BN, rows read in place from the rollout buffer) -> heads -> MultiCategorical
    -> PPO loss fwd+bwd -> MultiCategorical bwd -> heads bwd -> encoder bwd -> [NCCL all-reduce] -> clip -> Adam

The minibatch cursor, the Adam step count and the KL early stop (ppo_grid_obs.py:259-268) live on the device: the stop
is a sticky flag that freezes Adam, the BatchNorm running statistics and the log, so the sequence can be captured once
as a CUDA graph and replayed for every minibatch of an epoch; the host reads the flag back once per epoch.

Multi-GPU: one process per GPU, each with its own envs and rollout buffer.  Per minibatch ONE
logical all-reduce (mean) of the flat gradient bucket, whose leading slot carries the ranks' KL-stop votes; it is issued in
two pieces -- the Linear-layer slice (99.9 % of the bytes) on a side stream as soon as the Linear backward is done, under
the convolution backward, then the small conv slice + vote.  BatchNorm running statistics are averaged over the ranks at
the end of `train()` (each rank normalises its minibatches with its own batch statistics, as the reference's
per-minibatch semantics imply; only the running buffers used by the eval-mode rollout policy would drift).
"""
import ctypes
import io
import json
import os
import time
import zipfile
from collections import deque

import numpy as np
import torch

from . import _lib, ops
from . import dist as gdist
from .buffers import TensorRolloutBuffer_Grid_Obs
from .policy import ActorCriticPolicy_Train_Eval

CTL_CURSOR, CTL_STOP, CTL_ADAM_STEP, CTL_STOP_AT, CTL_LOGGED = 0, 1, 2, 3, 4


class _Logger:
    """Minimal stand-in for SB3's Logger (record / dump / dir); `name_to_value` keeps the last value of every key."""

    def __init__(self, folder=None):
        self.name_to_value = {}
        self.dir = folder

    def record(self, key, value, exclude=None):
        self.name_to_value[key] = value

    def dump(self, step=0):
        pass


class _Callback:
    """Normalises what `learn(callback=...)` may receive (on_policy_algorithm_grid_obs.py:158,185-186,219,249,296):
    an SB3 `BaseCallback`-like object (`init_callback / on_training_start / on_rollout_start / update_locals / on_step /
    on_rollout_end / on_training_end`), a list of those, a plain callable `f(locals) -> bool|None`, or None."""

    def __init__(self, cb, model):
        self.objs, self.fn = [], None
        if cb is None:
            return
        for c in (cb if isinstance(cb, (list, tuple)) else [cb]):
            if hasattr(c, "on_step") and hasattr(c, "update_locals"):
                if hasattr(c, "init_callback"):
                    c.init_callback(model)
                self.objs.append(c)
            elif callable(c):
                self.fn = c
            else:
                raise TypeError(f"unsupported callback {c!r}")

    def on_training_start(self, loc, glob):
        for c in self.objs:
            c.on_training_start(loc, glob)

    def on_rollout_start(self):
        for c in self.objs:
            c.on_rollout_start()

    def on_step(self, loc):
        ok = True
        for c in self.objs:
            c.update_locals(loc)
            ok = (c.on_step() is not False) and ok
        if self.fn is not None:
            ok = (self.fn(loc) is not False) and ok
        return ok

    def on_rollout_end(self):
        for c in self.objs:
            c.on_rollout_end()

    def on_training_end(self):
        for c in self.objs:
            c.on_training_end()


class PPO_Grid_Obs:
    def __init__(self, policy=ActorCriticPolicy_Train_Eval, env=None, learning_rate=3e-4, n_steps=2048, batch_size=64,
                 n_epochs=10, gamma=0.99, gae_lambda=0.95, clip_range=0.2, clip_range_vf=None, normalize_advantage=True,
                 ent_coef=0.0, vf_coef=0.5, max_grad_norm=0.5, use_sde=False, sde_sample_freq=-1, target_kl=None,
                 tensorboard_log=None, create_eval_env=False, policy_kwargs=None, verbose=0, seed=None, device="cuda",
                 _init_setup_model=True):
        if use_sde:
            raise ValueError("gennbv_b200 implements the GenNBV configuration (use_sde=False)")
        self.env = env
        self.device = torch.device(device)
        self.learning_rate, self.n_steps, self.batch_size, self.n_epochs = learning_rate, n_steps, batch_size, n_epochs
        self.gamma, self.gae_lambda = gamma, gae_lambda
        self.clip_range = clip_range if callable(clip_range) else (lambda _: clip_range)
        self.clip_range_vf = None if clip_range_vf is None else (clip_range_vf if callable(clip_range_vf) else (lambda _: clip_range_vf))
        self.normalize_advantage, self.ent_coef, self.vf_coef = normalize_advantage, ent_coef, vf_coef
        self.max_grad_norm, self.target_kl = max_grad_norm, target_kl
        self.policy_class, self.policy_kwargs = policy, dict(policy_kwargs or {})
        self.seed, self.verbose, self.tensorboard_log = seed, verbose, tensorboard_log
        self.bind_rollout_slots = True                       # observations are written straight into the buffer
        self.use_cuda_graph = os.environ.get("GNBV_PPO_GRAPH", "1") != "0"
        self.overlap_allreduce = True
        self.pg_coef = 10.0                                  # ppo_grid_obs.py:253 (`policy_loss * 10`)
        self.num_timesteps = self._n_updates = 0
        self._total_timesteps = 0
        self._current_progress_remaining = 1.0
        self._last_obs = self._last_episode_starts = None
        self.logger = _Logger(tensorboard_log)
        self.ep_info_buffer = deque(maxlen=100)              # base_class_grid_obs.py:436
        self.is_isaac_gym_env = True
        self.rank, self.world_size = gdist.world()
        self._graphs = {}
        self._overlap = None
        if _init_setup_model:
            self._setup_model()

    # on_policy_algorithm_grid_obs.py:102-126
    def _setup_model(self):
        if self.seed is not None:
            torch.manual_seed(self.seed)
            np.random.seed(self.seed)
            self.env.seed(self.seed)
        env = self.env
        self.observation_space, self.action_space, self.n_envs = env.observation_space, env.action_space, env.num_envs
        # every rank must run the same number of minibatches per epoch (one gradient all-reduce each): shards that differ
        # by an env (dist.shard_envs) would dead-lock the collectives, so refuse them up front
        gdist.require_equal(self.n_envs * self.n_steps, "n_envs * n_steps")
        self.rollout_buffer = TensorRolloutBuffer_Grid_Obs(self.n_steps, self.observation_space, self.action_space,
                                                           device=self.device, gamma=self.gamma,
                                                           gae_lambda=self.gae_lambda, n_envs=self.n_envs)
        lr = self.learning_rate
        self.lr_schedule = lr if callable(lr) else (lambda _: lr)
        self.policy = self.policy_class(self.observation_space, self.action_space, self.lr_schedule, device=self.device,
                                        **self.policy_kwargs)
        # action-sampling noise: seeded by the user's seed or, like the reference's stochastic default, by torch's RNG seed;
        # a different Philox key per rank so that env i of every rank does not draw the same noise
        base = self.seed if self.seed is not None else torch.initial_seed()
        self.policy.sample_seed = (int(base) * 1000003 + self.rank) & 0x7FFFFFFFFFFFFFFF
        L = _lib.lib()
        n = self.policy.flat_params.numel()
        self._exp_avg = torch.zeros(n, device=self.device)
        self._exp_avg_sq = torch.zeros(n, device=self.device)
        self._apply_ws = torch.zeros(L.gnbv_ppo_apply_workspace_bytes() // 4, device=self.device)
        self._ctl = torch.zeros(8, dtype=torch.int64, device=self.device)
        self._ctl[CTL_STOP_AT] = -1
        self._log = None
        self._mb_args = {}
        gdist.broadcast_state_(self.policy.flat_params, list(self.policy.buffers()))   # identical replicas

    @property
    def _adam_step(self):
        return int(self._ctl[CTL_ADAM_STEP])

    def get_env(self):
        return self.env

    # ---------------------------------------------------------------------------------------------------- rollouts
    def _setup_learn(self, total_timesteps=None, reset_num_timesteps=True):
        """base_class_grid_obs.py:425-477."""
        self.start_time = time.time()
        if reset_num_timesteps:
            self.ep_info_buffer = deque(maxlen=100)
            self.num_timesteps = 0
        elif total_timesteps is not None:
            total_timesteps += self.num_timesteps
        if total_timesteps is not None:
            self._total_timesteps = total_timesteps
        if reset_num_timesteps or self._last_obs is None:
            self._last_obs = self.env.reset()
            self._last_episode_starts = torch.ones(self.n_envs, dtype=torch.bool, device=self.device)
        # :472-475: the randomised counter lands on the wrapper, not on the env
        self.env.episode_length_buf = torch.randint_like(self.env.episode_length_buf, high=int(self.env.max_episode_length))
        return total_timesteps

    def collect_rollouts(self, env=None, callback=None, rollout_buffer=None, n_rollout_steps=None):
        """on_policy_algorithm_grid_obs.py:128-221."""
        env = self.env if env is None else env
        buf = self.rollout_buffer if rollout_buffer is None else rollout_buffer
        n_rollout_steps = self.n_steps if n_rollout_steps is None else n_rollout_steps
        callback = callback if isinstance(callback, _Callback) else _Callback(callback, self)
        assert self._last_obs is not None, "No previous observation was provided"
        self.policy.set_training_mode(False)
        buf.reset()
        n_steps = 0
        callback.on_rollout_start()
        # One encoder pass per env step instead of the reference's two: in eval mode with fixed weights
        # predict_values(new_obs) at step t and policy(obs) at step t+1 see the same observation,
        # so the features extracted for the bootstrap are reused for the next action.  Results are identical.
        with torch.no_grad():
            feats = self.policy.extract_features(self._last_obs)
        bind = getattr(env, "bind_next_observation", None) if self.bind_rollout_slots else None
        while n_steps < n_rollout_steps:
            actions, values, log_probs = self.policy.act_from_features(feats)
            if bind is not None and n_steps + 1 < min(n_rollout_steps, buf.buffer_size):
                bind(buf.observations[n_steps + 1])          # new_obs is born in the slot the next add() would copy it to
            new_obs, rewards, dones, infos = env.step(actions)
            self.num_timesteps += env.num_envs
            if not callback.on_step(locals()):
                return False
            # _update_info_buffer (base_class_grid_obs.py:491-493); the env reuses its statistics storage, so keep copies
            ep = infos.get("episode")
            self.ep_info_buffer.extend([{k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in ep.items()}
                                        if isinstance(ep, dict) else ep])
            n_steps += 1
            with torch.no_grad():
                feats = self.policy.extract_features(new_obs)
            new_values = self.policy.values_from_features(feats)
            # time-out bootstrap (:205-208); `[0]` picks env 0's value for every env -- reproduced, not fixed
            terminal_value = new_values[0]
            rewards += self.gamma * torch.squeeze(terminal_value * infos["time_outs"].unsqueeze(1).to(self.device), 1)
            buf.add(self._last_obs, actions, rewards, self._last_episode_starts, values, log_probs)
            self._last_obs, self._last_episode_starts = new_obs, dones
        values = new_values                                   # == policy.predict_values(new_obs) (:213-215)
        buf.compute_returns_and_advantage(last_values=values, dones=dones)
        callback.on_rollout_end()
        return True

    # ---------------------------------------------------------------------------------------------------- update
    def _minibatch_args(self, B, clip_range, clip_range_vf, rows_base=0):
        """The gnbv_ppo_minibatch argument block for batch size B (device pointers are stable for the algorithm's life)."""
        key = (B, float(clip_range), None if clip_range_vf is None else float(clip_range_vf), int(rows_base),
               float(self.lr_schedule(self._current_progress_remaining)))
        hit = self._mb_args.get(B)
        if hit is not None and hit["key"] == key:
            return hit
        pol, buf, L, dev = self.policy, self.rollout_buffer, _lib.lib(), self.device
        enc = pol.features_extractor
        keep = dict(key=key)
        keep["enc_p"] = enc._c_params()
        gp = _lib.EncoderGrads()
        keep["grad_views"] = pol.encoder_grad_views()
        for name, g in zip(_lib.EncoderGrads.FIELDS, keep["grad_views"]):
            setattr(gp, name, g.data_ptr())
        keep["enc_g"] = gp
        if hit is not None:                                  # same shapes: keep the workspaces (a captured graph points at them)
            keep["enc_ws"], keep["mb_ws"] = hit["enc_ws"], hit["mb_ws"]
        else:
            keep["enc_ws"] = torch.empty(L.gnbv_encoder_workspace_bytes(B, enc.grid_size, enc.state_dim, 1), dtype=torch.uint8, device=dev)
            keep["mb_ws"] = torch.empty(L.gnbv_ppo_minibatch_workspace_bytes(B, pol.num_logits, len(pol.nvec), pol.features_dim),
                                        dtype=torch.uint8, device=dev)
        a = _lib.PpoMinibatch()
        a.enc, a.enc_grads = ctypes.pointer(keep["enc_p"]), ctypes.pointer(keep["enc_g"])
        a.head_w, a.head_b = pol.head_w.data_ptr(), pol.head_b.data_ptr()
        a.head_w_grad, a.head_b_grad = pol.head_w_grad.data_ptr(), pol.head_b_grad.data_ptr()
        a.nvec, a.num_sub, a.feat_dim = pol._nvec_c, len(pol.nvec), pol.features_dim
        obs = buf.flat("observations")
        a.observations, a.obs_row_stride = obs.data_ptr(), obs.stride(0)
        a.actions, a.values, a.log_probs = buf.actions.data_ptr(), buf.values.data_ptr(), buf.log_probs.data_ptr()
        a.advantages, a.returns = buf.advantages.data_ptr(), buf.returns.data_ptr()
        a.storage_rows, a.rows_base = buf.storage_rows.data_ptr(), int(rows_base)
        a.batch, a.grid_size, a.state_dim = B, enc.grid_size, enc.state_dim
        a.normalize_advantage = int(self.normalize_advantage)
        a.clip_range, a.clip_range_vf = float(clip_range), -1.0 if clip_range_vf is None else float(clip_range_vf)
        a.ent_coef, a.vf_coef, a.pg_coef = float(self.ent_coef), float(self.vf_coef), float(self.pg_coef)
        a.target_kl = -1.0 if self.target_kl is None else float(self.target_kl)
        a.ctl, a.vote = self._ctl.data_ptr(), pol.grad_vote.data_ptr()
        a.log, a.log_capacity = self._log.data_ptr(), self._log.shape[0]
        a.enc_workspace, a.enc_workspace_bytes = keep["enc_ws"].data_ptr(), keep["enc_ws"].numel()
        a.mb_workspace, a.mb_workspace_bytes = keep["mb_ws"].data_ptr(), keep["mb_ws"].numel()
        keep["args"] = a
        self._mb_args[B] = keep
        self._graphs.pop(B, None)
        return keep

    def _launch_minibatch(self, keep):
        """Enqueues one minibatch update on the current stream (no host synchronisation)."""
        L, pol, a = _lib.lib(), self.policy, keep["args"]
        s = ops._stream()
        ws = self.world_size
        if ws > 1 and self.overlap_allreduce:
            if self._overlap is None:
                self._overlap = gdist.OverlappedGradAllreduce(pol.grad_bucket, 4 + pol.linear_slice_offset)
            _lib.check(L.gnbv_ppo_minibatch_grads(ctypes.byref(a), 1, s), "gnbv_ppo_minibatch_grads")
            self._overlap.start_linear()                             # Linear slice (99.9 % of the bytes) under the conv backward
            _lib.check(L.gnbv_ppo_minibatch_grads(ctypes.byref(a), 2, s), "gnbv_ppo_minibatch_grads")
            self._overlap.finish()                                   # votes + conv tensors (30 KB), then join
        else:
            _lib.check(L.gnbv_ppo_minibatch_grads(ctypes.byref(a), 3, s), "gnbv_ppo_minibatch_grads")
            gdist.allreduce_avg_(pol.grad_bucket)
        n = pol.flat_grads.numel()
        _lib.check(L.gnbv_ppo_minibatch_apply(ctypes.byref(a), pol.flat_params.data_ptr(), pol.flat_grads.data_ptr(),
                                              self._exp_avg.data_ptr(), self._exp_avg_sq.data_ptr(), n, float(self.max_grad_norm),
                                              float(self.lr_schedule(self._current_progress_remaining)), 0.9, 0.999,
                                              float(pol.optimizer_kwargs.get("eps", 1e-8)), 1.0, self._apply_ws.data_ptr(),
                                              ops._stream()), "gnbv_ppo_minibatch_apply")

    def _run_minibatch(self, keep):
        B = keep["args"].batch
        # With more than one rank the sequence contains two NCCL all-reduces and a side-stream fork / join.  Capturing them works
        # (measured 1.80 instead of 1.90 ms per update at 2 x B200) but is opt-in (GNBV_PPO_GRAPH_NCCL=1): one 2-rank test run
        # with captured collectives did not terminate, and a hung job costs more than the 5 %.
        if not self.use_cuda_graph or (self.world_size > 1 and os.environ.get("GNBV_PPO_GRAPH_NCCL", "0") != "1"):
            return self._launch_minibatch(keep)
        g = self._graphs.get(B)
        if g is None:
            # warm-up outside capture would advance the device state; capture directly (every launch is capture-safe:
            # no allocation, no synchronisation, shared-memory attributes were granted by earlier eager calls or are
            # granted before capture below)
            self._grant_smem_attributes(keep)
            g = torch.cuda.CUDAGraph()
            cap = torch.cuda.Stream()
            cap.wait_stream(torch.cuda.current_stream())
            with torch.cuda.graph(g, stream=cap, capture_error_mode="thread_local"):
                self._launch_minibatch(keep)
            torch.cuda.current_stream().wait_stream(cap)
            self._graphs[B] = g
        g.replay()

    def _grant_smem_attributes(self, keep):
        """cudaFuncSetAttribute is not allowed during stream capture: run the sequence once eagerly on a scratch copy of the
        control block so that every kernel has its dynamic shared-memory grant, then restore all touched state."""
        pol = self.policy
        saved = [t.clone() for t in (self._ctl, pol.flat_params, self._exp_avg, self._exp_avg_sq, self._log)]
        bufs = [b.clone() for b in pol.buffers()]
        self._launch_minibatch(keep)
        for t, s in zip((self._ctl, pol.flat_params, self._exp_avg, self._exp_avg_sq, self._log), saved):
            t.copy_(s)
        for b, s in zip(pol.buffers(), bufs):
            b.copy_(s)

    def train(self):
        """ppo_grid_obs.py:176-297."""
        t0 = time.time()
        self.policy.set_training_mode(True)
        clip_range = self.clip_range(self._current_progress_remaining)
        clip_range_vf = None if self.clip_range_vf is None else self.clip_range_vf(self._current_progress_remaining)
        buf = self.rollout_buffer
        assert buf.step == buf.buffer_size, ""
        total, B = buf.buffer_size * buf.n_envs, int(self.batch_size)
        n_full, tail = divmod(total, B)
        n_mb = n_full + (1 if tail else 0)
        cap = self.n_epochs * n_mb
        if self._log is None or self._log.shape[0] < cap:
            self._log = torch.zeros(cap, 8, device=self.device)
            self._mb_args.clear()
            self._graphs.clear()
        ctl = self._ctl
        ctl[CTL_STOP] = 0
        ctl[CTL_STOP_AT] = -1
        ctl[CTL_LOGGED] = 0
        full = self._minibatch_args(B, clip_range, clip_range_vf) if n_full else None
        # the truncated last minibatch (cursor = n_full) reads rows [n_full*B, total)
        part = self._minibatch_args(tail, clip_range, clip_range_vf, rows_base=n_full * (B - tail)) if tail else None
        stopped_epoch, per_epoch = None, []
        for epoch in range(self.n_epochs):
            ctl[CTL_CURSOR] = 0
            for _ in range(n_full):
                self._run_minibatch(full)
            if tail:
                self._run_minibatch(part)
            state = ctl.tolist()                 # ONE host read-back per epoch (the reference reads approx_kl per minibatch)
            per_epoch.append(state[CTL_LOGGED])
            if state[CTL_STOP]:
                stopped_epoch = epoch
                if self.verbose >= 1:
                    print(f"Early stopping at step {epoch} due to reaching max kl")
                break
        self._n_updates += self.n_epochs
        n_logged = per_epoch[-1] if per_epoch else 0
        sc = self._log[:n_logged].cpu().numpy() if n_logged else np.zeros((1, 8), np.float32)
        first_of_last = per_epoch[-2] if len(per_epoch) > 1 else 0
        if self.world_size > 1:
            self._sync_bn_buffers()
        y_pred, y_true = buf.values.flatten(), buf.returns.flatten()
        var_y = torch.var(y_true)
        explained_var = float("nan") if float(var_y) == 0 else float(1 - torch.var(y_true - y_pred) / var_y)
        rec = self.logger.record
        rec("train/entropy_loss", float(np.mean(sc[:, 3])))
        rec("train/policy_gradient_loss", float(np.mean(sc[:, 1])))
        rec("train/value_loss", float(np.mean(sc[:, 2])))
        rec("train/approx_kl", float(np.mean(sc[first_of_last:, 4])))       # the list is reset every epoch (:199)
        rec("train/clip_fraction", float(np.mean(sc[:, 5])))
        rec("train/loss", float(sc[-1, 0]))
        rec("train/explained_variance", explained_var)
        rec("train/n_updates", self._n_updates, exclude="tensorboard")
        rec("train/clip_range", clip_range)
        if clip_range_vf is not None:
            rec("train/clip_range_vf", clip_range_vf)
        rec("time/training", time.time() - t0)
        self._last_train = dict(minibatches_logged=n_logged, stopped_epoch=stopped_epoch, scalars=sc)

    def _sync_bn_buffers(self):
        """Average the BatchNorm running statistics over the ranks (one 64-float all-reduce per train()): the weights are
        kept identical by the gradient all-reduce, the running buffers follow each rank's own minibatches."""
        bufs = [b for b in self.policy.buffers() if b.dtype == torch.float32]
        flat = torch.cat([b.flatten() for b in bufs])
        gdist.allreduce_mean_(flat)
        o = 0
        for b in bufs:
            b.copy_(flat[o:o + b.numel()].view_as(b))
            o += b.numel()

    # ---------------------------------------------------------------------------------------------------- persistence
    def sync_optimizer_state(self):
        """Mirror the fused Adam state (flat exp_avg / exp_avg_sq arenas, step count) into `policy.optimizer.state`, so
        that `policy.optimizer.state_dict()` -- what save() writes as policy.optimizer.pth
        (on_policy_algorithm_grid_obs.py:300-303) -- reflects the training done by the fused path."""
        pol = self.policy
        step = float(self._adam_step)
        for p, (o, n) in zip(pol.arena_parameters(), pol._arena):
            st = pol.optimizer.state[p]
            st["step"] = torch.tensor(step)
            st["exp_avg"] = self._exp_avg[o:o + n].view_as(p)
            st["exp_avg_sq"] = self._exp_avg_sq[o:o + n].view_as(p)

    def get_parameters(self):
        """base_class_grid_obs.py:760-775: {"policy": state_dict, "policy.optimizer": state_dict}."""
        self.sync_optimizer_state()
        return {"policy": self.policy.state_dict(), "policy.optimizer": self.policy.optimizer.state_dict()}

    def save(self, path, exclude=None, include=None):
        """base_class_grid_obs.py:806-854 / save_util.py:289-330: a zip with `data` (JSON of the plain attributes),
        `policy.pth`, `policy.optimizer.pth`, `pytorch_variables.pth` and `_stable_baselines3_version` -- the member names
        and tensor keys the reference's `load_from_zip_file` / `set_parameters` read."""
        path = str(path)
        if not path.endswith(".zip"):
            path += ".zip"
        os.makedirs(os.path.dirname(os.path.abspath(path)) or ".", exist_ok=True)
        params = self.get_parameters()
        data = {k: getattr(self, k) for k in ("n_steps", "batch_size", "n_epochs", "gamma", "gae_lambda", "normalize_advantage",
                                              "ent_coef", "vf_coef", "max_grad_norm", "target_kl", "num_timesteps",
                                              "_n_updates", "_total_timesteps", "seed", "n_envs", "pg_coef")}
        data["learning_rate"] = self.lr_schedule(1.0)
        data["clip_range"] = self.clip_range(1.0)
        data["clip_range_vf"] = None if self.clip_range_vf is None else self.clip_range_vf(1.0)
        with zipfile.ZipFile(path, "w") as z:
            z.writestr("data", json.dumps(data, indent=4))
            for name, sd in params.items():
                f = io.BytesIO()
                torch.save({k: (v.detach().cpu() if isinstance(v, torch.Tensor) else v) for k, v in sd.items()}
                           if name == "policy" else _cpu_tree(sd), f)
                z.writestr(name + ".pth", f.getvalue())
            f = io.BytesIO()
            torch.save({}, f)
            z.writestr("pytorch_variables.pth", f.getvalue())
            z.writestr("_stable_baselines3_version", "1.6.0")
        return path

    def set_parameters(self, load_path_or_dict, exact_match=True, device="auto"):
        """base_class_grid_obs.py:600-690: load `policy` (+ `policy.optimizer`) from a dict or from a zip written by
        save() / by the reference's save(); the Adam moments are copied into the fused flat arenas."""
        if isinstance(load_path_or_dict, dict):
            params = load_path_or_dict
        else:
            path = str(load_path_or_dict)
            if not os.path.exists(path) and os.path.exists(path + ".zip"):
                path += ".zip"
            params = {}
            with zipfile.ZipFile(path) as z:
                for name in ("policy", "policy.optimizer"):
                    if name + ".pth" in z.namelist():
                        params[name] = torch.load(io.BytesIO(z.read(name + ".pth")), map_location="cpu", weights_only=False)
        if "policy" not in params and exact_match:
            raise ValueError("set_parameters: no `policy` entry")
        if "policy" in params:
            self.policy.load_state_dict(params["policy"], strict=exact_match)
        opt = params.get("policy.optimizer")
        if opt is not None and opt.get("state"):
            pol = self.policy
            step = 0
            for i, (p, (o, n)) in enumerate(zip(pol.optimizer_parameter_order(), pol.optimizer_arena_order())):
                st = opt["state"].get(i)
                if st is None:
                    continue
                self._exp_avg[o:o + n].copy_(st["exp_avg"].reshape(-1))
                self._exp_avg_sq[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
                step = max(step, int(float(st["step"])))
            self._ctl[CTL_ADAM_STEP] = step
        elif exact_match and "policy.optimizer" not in params and not isinstance(load_path_or_dict, dict):
            raise ValueError("set_parameters: no `policy.optimizer` entry")
        self._graphs.clear()

    @classmethod
    def load(cls, path, env=None, device="cuda", **kwargs):
        """Rebuild an algorithm from save()'s zip (hyper-parameters from `data`, then set_parameters)."""
        p = str(path)
        if not os.path.exists(p) and os.path.exists(p + ".zip"):
            p += ".zip"
        with zipfile.ZipFile(p) as z:
            data = json.loads(z.read("data"))
        args = {k: data[k] for k in ("learning_rate", "n_steps", "batch_size", "n_epochs", "gamma", "gae_lambda", "clip_range",
                                     "clip_range_vf", "normalize_advantage", "ent_coef", "vf_coef", "max_grad_norm",
                                     "target_kl", "seed")}
        args.update(kwargs)
        model = cls(env=env, device=device, **args)
        model.num_timesteps, model._n_updates = data["num_timesteps"], data["_n_updates"]
        model.set_parameters(p)
        return model

    def predict(self, observation, state=None, episode_start=None, deterministic=False):
        """base_class_grid_obs.py:578-598."""
        return self.policy.predict(observation, state, episode_start, deterministic)

    def learn(self, total_timesteps, callback=None, log_interval=1, eval_env=None, eval_freq=-1, n_eval_episodes=5,
              tb_log_name="PPO", eval_log_path=None, reset_num_timesteps=True):
        """on_policy_algorithm_grid_obs.py:230-298."""
        iteration = 0
        total_timesteps = self._setup_learn(total_timesteps, reset_num_timesteps)
        callback = _Callback(callback, self)
        callback.on_training_start(locals(), globals())
        while self.num_timesteps < total_timesteps:
            t0 = time.time()
            if not self.collect_rollouts(self.env, callback, self.rollout_buffer, n_rollout_steps=self.n_steps):
                break
            iteration += 1
            self._current_progress_remaining = 1.0 - float(self.num_timesteps) / float(total_timesteps)
            if log_interval is not None and iteration % log_interval == 0:
                dt = time.time() - t0
                rec = self.logger.record
                rec("time/iterations", iteration, exclude="tensorboard")
                if len(self.ep_info_buffer) > 0 and isinstance(self.ep_info_buffer[0], dict):
                    for key in self.ep_info_buffer[0]:                        # 'rollout/rew_*', 'rollout/episode_*' (:269-279)
                        vals = torch.stack([torch.as_tensor(e[key], dtype=torch.float32, device=self.device).reshape(())
                                            for e in self.ep_info_buffer])
                        rec("rollout/{}".format(key), np.round(vals.mean().cpu().numpy(), 4))
                rec("time/fps", int(self.rollout_buffer.buffer_size * self.rollout_buffer.n_envs / dt))
                rec("time/time_elapsed", int(time.time() - self.start_time), exclude="tensorboard")
                rec("time/total_timesteps", self.num_timesteps, exclude="tensorboard")
                rec("time/rollout", dt, exclude="tensorboard")
                self.logger.dump(step=self.num_timesteps)
            self.train()
        callback.on_training_end()
        return self


def _cpu_tree(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().clone()
    if isinstance(x, dict):
        return {k: _cpu_tree(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_cpu_tree(v) for v in x)
    return x

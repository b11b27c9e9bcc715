"""`PPO_Grid_Obs` -- the GenNBV PPO variant (stable_baselines3/ppo/ppo_grid_obs.py +
stable_baselines3/common/on_policy_algorithm_grid_obs.py) on the kernels of libgennbv_b200.

`collect_rollouts()` keeps the reference's control flow (policy forward in eval mode, env.step, time-out bootstrap
with the extra predict_values pass, buffer.add) and `train()` keeps its arithmetic and logged scalars, but a minibatch
update is a fixed sequence of kernel launches on flat parameter / gradient arenas with no autograd graph:

    encoder forward (batch-stat BN, rows read in place from the rollout buffer) -> heads -> MultiCategorical
    -> PPO loss fwd+bwd -> MultiCategorical bwd -> heads bwd -> encoder bwd -> [NCCL all-reduce of the flat gradient]
    -> global-norm clip -> Adam.

Multi-GPU: one process per GPU, each with its own envs and rollout buffer; the only data-path
collective is the all-reduce (mean) of the flat gradient before the clip, plus a MAX all-reduce of approx_kl so that
the early stop is decided identically on every rank.
"""
import ctypes
import time

import numpy as np
import torch

from . import _lib, ops
from . import dist as gdist
from .buffers import TensorRolloutBuffer_Grid_Obs
from .policy import ActorCriticPolicy_Train_Eval


class _Logger:
    def __init__(self):
        self.name_to_value = {}

    def record(self, key, value, exclude=None):
        self.name_to_value[key] = value

    def dump(self, step=0):
        pass


class PPO_Grid_Obs:
    def __init__(self, policy=ActorCriticPolicy_Train_Eval, env=None, learning_rate=3e-4, n_steps=2048, batch_size=64,
                 n_epochs=10, gamma=0.99, gae_lambda=0.95, clip_range=0.2, clip_range_vf=None, normalize_advantage=True,
                 ent_coef=0.0, vf_coef=0.5, max_grad_norm=0.5, target_kl=None, policy_kwargs=None, seed=None,
                 device="cuda", verbose=0, tensorboard_log=None, create_eval_env=False, _init_setup_model=True):
        self.env = env
        self.device = torch.device(device)
        self.learning_rate, self.n_steps, self.batch_size, self.n_epochs = learning_rate, n_steps, batch_size, n_epochs
        self.gamma, self.gae_lambda = gamma, gae_lambda
        self.clip_range = clip_range if callable(clip_range) else (lambda _: clip_range)
        self.clip_range_vf = None if clip_range_vf is None else (clip_range_vf if callable(clip_range_vf) else (lambda _: clip_range_vf))
        self.normalize_advantage, self.ent_coef, self.vf_coef = normalize_advantage, ent_coef, vf_coef
        self.max_grad_norm, self.target_kl = max_grad_norm, target_kl
        self.policy_class, self.policy_kwargs = policy, dict(policy_kwargs or {})
        self.seed, self.verbose = seed, verbose
        self.bind_rollout_slots = True                       # observations are written straight into the buffer
        self.pg_coef = 10.0                                  # ppo_grid_obs.py:253 (`policy_loss * 10`)
        self.num_timesteps = self._n_updates = 0
        self._current_progress_remaining = 1.0
        self._last_obs = self._last_episode_starts = None
        self.logger = _Logger()
        self.ep_info_buffer = []
        self.is_isaac_gym_env = True
        self.world_size = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
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
        self.rollout_buffer = TensorRolloutBuffer_Grid_Obs(self.n_steps, self.observation_space, self.action_space,
                                                           device=self.device, gamma=self.gamma,
                                                           gae_lambda=self.gae_lambda, n_envs=self.n_envs)
        lr = self.learning_rate
        self.lr_schedule = lr if callable(lr) else (lambda _: lr)
        self.policy = self.policy_class(self.observation_space, self.action_space, self.lr_schedule, device=self.device,
                                        **self.policy_kwargs)
        if self.seed is not None:
            self.policy.sample_seed = self.seed
        n = self.policy.flat_params.numel()
        self._exp_avg = torch.zeros(n, device=self.device)
        self._exp_avg_sq = torch.zeros(n, device=self.device)
        self._adam_step = 0
        self._clip_ws = torch.zeros(_lib.lib().gnbv_clip_adam_workspace_bytes() // 4, device=self.device)
        self._scalars = torch.zeros(8, device=self.device)
        gdist.broadcast_state_(self.policy.flat_params, list(self.policy.buffers()))   # identical replicas

    # ---------------------------------------------------------------------------------------------------- rollouts
    def _setup_learn(self):
        self._last_obs = self.env.reset()
        self._last_episode_starts = torch.ones(self.n_envs, dtype=torch.bool, device=self.device)
        # base_class_grid_obs.py:472-475: the randomised counter lands on the wrapper, not on the env
        self.env.episode_length_buf = torch.randint_like(self.env.episode_length_buf, high=int(self.env.max_episode_length))

    def collect_rollouts(self, env=None, callback=None, rollout_buffer=None, n_rollout_steps=None):
        """on_policy_algorithm_grid_obs.py:128-221."""
        env = self.env if env is None else env
        buf = self.rollout_buffer if rollout_buffer is None else rollout_buffer
        n_rollout_steps = self.n_steps if n_rollout_steps is None else n_rollout_steps
        assert self._last_obs is not None, "No previous observation was provided"
        self.policy.set_training_mode(False)
        buf.reset()
        n_steps = 0
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
            if callback is not None and callback(locals()) is False:
                return False
            self.ep_info_buffer.append(infos.get("episode"))
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
        return True

    # ---------------------------------------------------------------------------------------------------- update
    def _minibatch_update(self, rows, clip_range, clip_range_vf):
        """One optimizer step on the rollout-buffer rows `rows` ([B] i64 on the device)."""
        pol, buf, L, s = self.policy, self.rollout_buffer, _lib.lib(), ops._stream()
        enc = pol.features_extractor
        B, A, F = rows.shape[0], pol.num_logits, pol.features_dim
        obs = buf.flat("observations")
        w = self._mb_ws(B)
        # gather the scalar columns of the minibatch (tiny) -- torch indexing is plumbing here
        actions = buf.flat("actions")[rows].long().contiguous()
        old_v = buf.flat("values")[rows].flatten().contiguous()
        old_lp = buf.flat("log_probs")[rows].flatten().contiguous()
        adv = buf.flat("advantages")[rows].flatten().contiguous()
        ret = buf.flat("returns")[rows].flatten().contiguous()
        ws = enc._workspace(B, obs.device, True)
        enc._run_forward(obs, need_bwd=True, row_index=rows, training=True, feats=w["feats"])
        _lib.check(L.gnbv_policy_heads_forward(w["feats"].data_ptr(), pol.head_w.data_ptr(), pol.head_b.data_ptr(),
                                               w["out"].data_ptr(), B, F, A + 1, s), "gnbv_policy_heads_forward")
        _lib.check(L.gnbv_multicategorical_evaluate(w["out"].data_ptr(), A + 1, pol._nvec_c, len(pol.nvec), actions.data_ptr(),
                                                    w["lp"].data_ptr(), w["ent"].data_ptr(), B, s), "multicategorical_evaluate")
        w["values"].copy_(w["out"][:, A])
        _lib.check(L.gnbv_ppo_loss(w["lp"].data_ptr(), w["ent"].data_ptr(), w["values"].data_ptr(), old_v.data_ptr(),
                                   old_lp.data_ptr(), adv.data_ptr(), ret.data_ptr(), B, float(clip_range),
                                   -1.0 if clip_range_vf is None else float(clip_range_vf), float(self.ent_coef),
                                   float(self.vf_coef), float(self.pg_coef), int(self.normalize_advantage),
                                   self._scalars.data_ptr(), w["g_lp"].data_ptr(), w["g_ent"].data_ptr(), w["g_v"].data_ptr(), s),
                   "gnbv_ppo_loss")
        return w, ws, actions

    def _minibatch_backward_and_step(self, rows, w, ws, actions):
        pol, buf, L, s = self.policy, self.rollout_buffer, _lib.lib(), ops._stream()
        enc = pol.features_extractor
        B, A, F = rows.shape[0], pol.num_logits, pol.features_dim
        _lib.check(L.gnbv_multicategorical_backward(w["out"].data_ptr(), A + 1, pol._nvec_c, len(pol.nvec), actions.data_ptr(),
                                                    w["g_lp"].data_ptr(), w["g_ent"].data_ptr(), w["dout"].data_ptr(), A + 1, B, s),
                   "gnbv_multicategorical_backward")
        w["dout"][:, A] = w["g_v"]
        ops.sgemm(w["dout"], (A + 1, 1), pol.head_w, (F, 1), w["dfeat"], B, F, A + 1)
        ops.sgemm(w["dout"], (1, A + 1), w["feats"], (F, 1), pol.head_w_grad, A + 1, F, B)
        torch.sum(w["dout"], dim=0, out=pol.head_b_grad)
        enc._run_backward(buf.flat("observations"), w["feats"], w["dfeat"], B, True, ws, row_index=rows, grads=self._enc_grads)
        n = pol.flat_grads.numel()
        gdist.allreduce_mean_(pol.flat_grads)                        # NCCL over NVLink, one flat bucket, before the clip
        _lib.check(L.gnbv_grad_norm(pol.flat_grads.data_ptr(), n, float(self.max_grad_norm), self._clip_ws.data_ptr(), s),
                   "gnbv_grad_norm")
        self._adam_step += 1
        _lib.check(L.gnbv_adam_step(pol.flat_params.data_ptr(), pol.flat_grads.data_ptr(), self._exp_avg.data_ptr(),
                                    self._exp_avg_sq.data_ptr(), n, self._clip_ws.data_ptr(), float(self.lr_schedule(
                                        self._current_progress_remaining)), 0.9, 0.999,
                                    float(pol.optimizer_kwargs.get("eps", 1e-8)), self._adam_step, 1.0, s),
                   "gnbv_adam_step")

    def sync_optimizer_state(self):
        """Mirror the fused Adam state (flat exp_avg / exp_avg_sq arenas, step count) into `policy.optimizer.state`, so
        that `policy.optimizer.state_dict()` -- what SB3's save() pickles as policy.optimizer.pth
        (on_policy_algorithm_grid_obs.py:300-303) -- reflects the training done by the fused path."""
        pol = self.policy
        order = pol.features_extractor._param_list() + [pol.action_net.weight, pol.value_net.weight, pol.action_net.bias,
                                                        pol.value_net.bias]
        for p, (o, n) in zip(order, pol._arena):
            st = pol.optimizer.state[p]
            st["step"] = torch.tensor(float(self._adam_step))
            st["exp_avg"] = self._exp_avg[o:o + n].view_as(p)
            st["exp_avg_sq"] = self._exp_avg_sq[o:o + n].view_as(p)

    def _mb_ws(self, B):
        if getattr(self, "_mb", None) is None or self._mb["feats"].shape[0] != B:
            pol, dev = self.policy, self.device
            A, F = pol.num_logits, pol.features_dim
            z = lambda *shape: torch.zeros(*shape, device=dev)
            self._mb = dict(feats=z(B, F), out=z(B, A + 1), lp=z(B), ent=z(B), values=z(B), g_lp=z(B), g_ent=z(B), g_v=z(B),
                            dout=z(B, A + 1), dfeat=z(B, F))
            self._enc_grads = pol.encoder_grad_views()
        return self._mb

    def train(self):
        """ppo_grid_obs.py:176-297."""
        t0 = time.time()
        self.policy.set_training_mode(True)
        clip_range = self.clip_range(self._current_progress_remaining)
        clip_range_vf = None if self.clip_range_vf is None else self.clip_range_vf(self._current_progress_remaining)
        log = []                       # per-minibatch device scalars, read back once at the end
        continue_training = True
        for epoch in range(self.n_epochs):
            for rows in self.rollout_buffer.minibatch_rows(self.batch_size):
                w, ws, actions = self._minibatch_update(rows, clip_range, clip_range_vf)
                log.append(self._scalars.clone())
                # one host read per minibatch, as the reference (:259-268); MAX over ranks keeps the ranks in lock-step
                if gdist.should_stop(self._scalars[4:5], self.target_kl):
                    continue_training = False
                    if self.verbose >= 1:
                        print(f"Early stopping at step {epoch} due to reaching max kl")
                    break
                self._minibatch_backward_and_step(rows, w, ws, actions)
            if not continue_training:
                break
        self._n_updates += self.n_epochs
        sc = torch.stack(log).cpu().numpy() if log else np.zeros((1, 8), np.float32)
        buf = self.rollout_buffer
        y_pred, y_true = buf.values.flatten(), buf.returns.flatten()
        var_y = torch.var(y_true)
        explained_var = float("nan") if float(var_y) == 0 else float(1 - torch.var(y_true - y_pred) / var_y)
        rec = self.logger.record
        rec("train/entropy_loss", float(np.mean(sc[:, 3])))
        rec("train/policy_gradient_loss", float(np.mean(sc[:, 1])))
        rec("train/value_loss", float(np.mean(sc[:, 2])))
        rec("train/approx_kl", float(np.mean(sc[:, 4])))
        rec("train/clip_fraction", float(np.mean(sc[:, 5])))
        rec("train/loss", float(sc[-1, 0]))
        rec("train/explained_variance", explained_var)
        rec("train/n_updates", self._n_updates, exclude="tensorboard")
        rec("train/clip_range", clip_range)
        if clip_range_vf is not None:
            rec("train/clip_range_vf", clip_range_vf)
        rec("time/training", time.time() - t0)

    def predict(self, observation, state=None, episode_start=None, deterministic=False):
        """base_class_grid_obs.py:578-598."""
        return self.policy.predict(observation, state, episode_start, deterministic)

    def learn(self, total_timesteps, callback=None, log_interval=1, **unused):
        """on_policy_algorithm_grid_obs.py:230-298."""
        self._setup_learn()
        iteration = 0
        self._total_timesteps = total_timesteps
        while self.num_timesteps < total_timesteps:
            t0 = time.time()
            if not self.collect_rollouts(callback=callback):
                break
            iteration += 1
            self._current_progress_remaining = 1.0 - float(self.num_timesteps) / float(total_timesteps)
            self.logger.record("time/rollout", time.time() - t0)
            self.train()
        return self

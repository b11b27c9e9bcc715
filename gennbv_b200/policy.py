"""`ActorCriticPolicy_Train_Eval` -- the actor-critic wrapper of the reference
(stable_baselines3/common/policies.py:797-1090, with `net_arch=[]` as train_gennbv.py:152 sets it: both heads sit
directly on the encoder output) on top of the CUDA kernels of libgennbv_b200.

Kept: `forward(obs, deterministic) -> (actions, values, log_prob)`, `evaluate_actions(obs, actions) ->
(values, log_prob, entropy)`, `predict_values(obs)`, `set_training_mode`, `.optimizer` (Adam, eps 1e-5, :851-855,997),
the orthogonal initialisation rule (:403-410,980-995: gain sqrt(2) on the encoder's Linear layers, 0.01 on
action_net, 1 on value_net; Conv3d keeps torch's default) and the `state_dict()` key names.

All parameters live in ONE flat fp32 arena (and all gradients in another): the fused clip+Adam kernels and the
NCCL all-reduce of the data-parallel build operate on the flat buffers; the nn.Parameters are views into them.
"""
import ctypes
from functools import partial

import numpy as np
import torch
from torch import nn

from . import _lib, ops
from .network import Hybrid_Encoder


def _nvec_array(nvec):
    arr = (ctypes.c_int * len(nvec))(*[int(v) for v in nvec])
    return arr


class _HeadsFn(torch.autograd.Function):
    """action_net + value_net + MultiCategorical log_prob / entropy (policies.py:1052-1068; distributions.py:323-337)."""

    @staticmethod
    def forward(ctx, feats, action_w, action_b, value_w, value_b, actions, policy):
        # the four head tensors are adjacent views of one [A+1, F] / [A+1] storage (policy.head_w / head_b)
        head_w, head_b = policy.head_w, policy.head_b
        B, A = feats.shape[0], policy.num_logits
        L = _lib.lib()
        out = torch.empty(B, A + 1, device=feats.device)
        _lib.check(L.gnbv_policy_heads_forward(feats.data_ptr(), head_w.data_ptr(), head_b.data_ptr(), out.data_ptr(), B,
                                               feats.shape[1], A + 1, ops._stream()), "gnbv_policy_heads_forward")
        lp = torch.empty(B, device=feats.device)
        ent = torch.empty(B, device=feats.device)
        _lib.check(L.gnbv_multicategorical_evaluate(out.data_ptr(), A + 1, policy._nvec_c, len(policy.nvec),
                                                    actions.data_ptr(), lp.data_ptr(), ent.data_ptr(), B, ops._stream()),
                   "gnbv_multicategorical_evaluate")
        ctx.policy = policy
        ctx.save_for_backward(feats, head_w, out, actions)
        return out[:, A:A + 1], lp, ent

    @staticmethod
    def backward(ctx, dvalues, dlp, dent):
        policy = ctx.policy
        feats, head_w, out, actions = ctx.saved_tensors
        B, A, F = feats.shape[0], policy.num_logits, feats.shape[1]
        L, s = _lib.lib(), ops._stream()
        dout = torch.empty(B, A + 1, device=feats.device)
        dlp = torch.zeros(B, device=feats.device) if dlp is None else dlp.contiguous()
        dent = torch.zeros(B, device=feats.device) if dent is None else dent.contiguous()
        _lib.check(L.gnbv_multicategorical_backward(out.data_ptr(), A + 1, policy._nvec_c, len(policy.nvec), actions.data_ptr(),
                                                    dlp.data_ptr(), dent.data_ptr(), dout.data_ptr(), A + 1, B, s),
                   "gnbv_multicategorical_backward")
        dout[:, A] = 0 if dvalues is None else dvalues.reshape(B)
        dfeat, dW = torch.empty(B, F, device=feats.device), torch.empty_like(head_w)
        ops.sgemm(dout, (A + 1, 1), head_w, (F, 1), dfeat, B, F, A + 1)                    # dfeat = dout W
        ops.sgemm(dout, (1, A + 1), feats, (F, 1), dW, A + 1, F, B)                        # dW = dout^T feats
        db = dout.sum(dim=0)
        return dfeat, dW[:A], db[:A], dW[A:], db[A:], None, None


class ActorCriticPolicy_Train_Eval(nn.Module):
    def __init__(self, observation_space, action_space, lr_schedule, net_arch=None, activation_fn=nn.Tanh, ortho_init=True,
                 features_extractor_class=Hybrid_Encoder, features_extractor_kwargs=None, optimizer_class=torch.optim.Adam,
                 optimizer_kwargs=None, device="cuda", **unused):
        super().__init__()
        if net_arch not in (None, []):
            raise ValueError("gennbv_b200 implements the GenNBV configuration net_arch=[] (train_gennbv.py:152)")
        if optimizer_kwargs is None:
            optimizer_kwargs = {}
            if optimizer_class == torch.optim.Adam:
                optimizer_kwargs["eps"] = 1e-5                              # policies.py:851-855
        self.observation_space, self.action_space = observation_space, action_space
        self.nvec = [int(v) for v in action_space.nvec]
        self.num_logits = int(sum(self.nvec))
        self._nvec_c = _nvec_array(self.nvec)
        self.ortho_init = ortho_init
        self.optimizer_class, self.optimizer_kwargs = optimizer_class, optimizer_kwargs
        self.features_extractor = features_extractor_class(observation_space, **(features_extractor_kwargs or {}))
        self.features_dim = self.features_extractor.features_dim
        self.action_net = nn.Linear(self.features_dim, self.num_logits)    # distributions.py:320
        self.value_net = nn.Linear(self.features_dim, 1)                   # policies.py:994
        if ortho_init:
            self.features_extractor.apply(partial(self.init_weights, gain=np.sqrt(2)))
            self.action_net.apply(partial(self.init_weights, gain=0.01))
            self.value_net.apply(partial(self.init_weights, gain=1))
        self.to(device)
        self._build_arena()
        self.optimizer = optimizer_class(self.parameters(), lr=lr_schedule(1), **optimizer_kwargs)
        self._sample_calls = 0
        self.sample_seed = 0

    @staticmethod
    def init_weights(module, gain=1.0):
        """policies.py:403-410: only Linear / Conv2d are re-initialised (Conv3d keeps torch's default)."""
        if isinstance(module, (nn.Linear, nn.Conv2d)):
            nn.init.orthogonal_(module.weight, gain=gain)
            if module.bias is not None:
                module.bias.data.fill_(0.0)

    # ---- flat parameter / gradient arenas ---------------------------------------------------------------------------
    def _build_arena(self):
        enc = self.features_extractor
        order = enc._param_list() + [self.action_net.weight, self.value_net.weight, self.action_net.bias, self.value_net.bias]
        assert len(order) == len(list(self.parameters()))
        sizes = [p.numel() for p in order]
        offs = np.concatenate([[0], np.cumsum([(n + 3) // 4 * 4 for n in sizes])])    # 16-byte aligned slices
        dev = order[0].device
        self.flat_params = torch.zeros(int(offs[-1]), device=dev)
        # gradient bucket = [vote slot (4 floats, 16 B) | flat gradients]: the data-parallel all-reduce carries the KL-stop
        # votes of the ranks in the same (small, conv-side) message as the gradients (csrc/ppo_update.cu)
        self.grad_bucket = torch.zeros(4 + int(offs[-1]), device=dev)
        self.grad_vote = self.grad_bucket[0:1]
        self.flat_grads = self.grad_bucket[4:]
        self._arena_order = order
        self._arena = []
        for p, o, n in zip(order, offs[:-1], sizes):
            view = self.flat_params[o:o + n].view_as(p)
            view.copy_(p.data)
            p.data = view
            p.grad = self.flat_grads[o:o + n].view_as(p)
            self._arena.append((int(o), int(n)))
        A, F = self.num_logits, self.features_dim
        o_w = self._arena[-4][0]
        self.head_w = self.flat_params[o_w:o_w + (A + 1) * F].view(A + 1, F)       # action_net.weight | value_net.weight
        o_b = self._arena[-2][0]
        assert self._arena[-3][0] == o_w + A * F and A % 4 == 0, "head weights must be adjacent"
        self.head_b = self.flat_params[o_b:o_b + A + 1]
        assert self._arena[-1][0] == o_b + A
        self.head_w_grad = self.flat_grads[o_w:o_w + (A + 1) * F].view(A + 1, F)
        self.head_b_grad = self.flat_grads[o_b:o_b + A + 1]

    def arena_parameters(self):
        """The parameters in flat-arena order (`_arena[i]` = (offset, numel) of the i-th)."""
        return list(self._arena_order)

    def optimizer_parameter_order(self):
        return list(self.parameters())                      # the order torch.optim indexes `state` by

    def optimizer_arena_order(self):
        where = {id(p): on for p, on in zip(self._arena_order, self._arena)}
        return [where[id(p)] for p in self.parameters()]

    @property
    def linear_slice_offset(self):
        """First element of `flat_grads` that belongs to a Linear layer (grid_fc.weight): everything from here on is final
        after the Linear phase of the backward, the conv tensors before it only after the conv phase."""
        return self._arena[8][0]

    def encoder_grad_views(self):
        """Views of the flat gradient arena for the encoder's 16 tensors, in gnbv_encoder_grads order (independent of
        whatever `p.grad` currently points to)."""
        enc_params = self.features_extractor._param_list()
        return [self.flat_grads[o:o + n].view_as(p) for p, (o, n) in zip(enc_params, self._arena)]

    def set_training_mode(self, mode):
        self.train(mode)

    # ---- reference API --------------------------------------------------------------------------------------------------
    def extract_features(self, obs):
        return self.features_extractor(obs.float())                                # preprocessing.py:101-104 (Box -> float)

    def _heads(self, feats):
        B, A = feats.shape[0], self.num_logits
        out = torch.empty(B, A + 1, device=feats.device)
        _lib.check(_lib.lib().gnbv_policy_heads_forward(feats.data_ptr(), self.head_w.data_ptr(), self.head_b.data_ptr(),
                                                        out.data_ptr(), B, feats.shape[1], A + 1, ops._stream()),
                   "gnbv_policy_heads_forward")
        return out

    def forward(self, obs, deterministic=False):
        """policies.py:999-1015 -> actions [N,6] i64, values [N,1], log_prob [N]."""
        with torch.no_grad():
            return self.act_from_features(self.extract_features(obs), deterministic)

    def values_from_features(self, feats):
        """value_net on already-extracted features (rollout-time reuse of one encoder pass per step)."""
        with torch.no_grad():
            return self._heads(feats)[:, self.num_logits:]

    def act_from_features(self, feats, deterministic=False):
        with torch.no_grad():
            out = self._heads(feats)
            B, A = out.shape[0], self.num_logits
            actions = torch.empty(B, len(self.nvec), dtype=torch.int64, device=out.device)
            lp = torch.empty(B, device=out.device)
            self._sample_calls += 1
            _lib.check(_lib.lib().gnbv_multicategorical_sample(out.data_ptr(), A + 1, self._nvec_c, len(self.nvec),
                                                               int(self.sample_seed), int(self._sample_calls) * 4096,
                                                               int(bool(deterministic)), actions.data_ptr(), lp.data_ptr(), B,
                                                               ops._stream()), "gnbv_multicategorical_sample")
        return actions, out[:, A:A + 1], lp

    def evaluate_actions(self, obs, actions):
        """policies.py:1052-1068 -> values [N,1], log_prob [N], entropy [N] (autograd-capable)."""
        feats = self.extract_features(obs)
        actions = actions.long().contiguous()
        if torch.is_grad_enabled():
            return _HeadsFn.apply(feats, self.action_net.weight, self.action_net.bias, self.value_net.weight,
                                  self.value_net.bias, actions, self)
        return _HeadsFn.forward(_NoCtx(), feats, None, None, None, None, actions, self)

    def predict_values(self, obs):
        """policies.py:1081-1090."""
        with torch.no_grad():
            return self._heads(self.extract_features(obs))[:, self.num_logits:]

    def predict(self, observation, state=None, episode_start=None, deterministic=False):
        """policies.py:425-477 -> (actions, state).  Eval mode; actions stay on the device as an int64 tensor (the
        reference converts them to numpy and the eval env converts them straight back, env_eval_gennbv.py:138-141)."""
        self.set_training_mode(False)
        return self.forward(observation, deterministic=deterministic)[0], state


class _NoCtx:
    def save_for_backward(self, *a):
        pass

"""`Hybrid_Encoder` -- drop-in for gennbv/network/hybrid_encoder.py::Hybrid_Encoder whose forward and backward are
the hand-written kernels of libgennbv_b200 (gnbv_encoder_forward / gnbv_encoder_backward).

Same constructor (`observation_space, encoder_param, net_param, visual_input_shape, state_input_shape`, including
the pop of `net_param["append_hidden_shapes"][-1]`, hybrid_encoder.py:27-28), same `features_dim`, same sub-module
names, hence the same `state_dict()` keys as the released checkpoint.  The torch sub-modules are
parameter containers only -- their own forward is never called -- and they are created in the reference's order so
that a given torch seed yields the same initial weights.

Generalisation: the reference hard-codes a 20^3 grid (`8000`, `20`, `1024`; hybrid_encoder.py:40,83-84).  Here the
grid size is derived from the observation length: D = state_dim + G^3 + rgb_dim.

`semantic_branch=True` (off by default) adds the paper's 2-D branch over the k = 2 grayscale frames that
the observation already carries behind the grid: `naive_encoder_rgb` (Conv2d(2,16,3,s2)+ReLU, Conv2d(16,16,3,s2)+ReLU),
`output_layer_rgb` (Linear(3600,256)+ReLU), and `output_layer` becomes Linear(768,256) -- the `[num_env, 256*3]` of
hybrid_encoder.py:89.  The released forward never reads the frames, so this branch has no reference to pin against
(parity unpinned; checked against torch autograd of the same layers).
"""
import ctypes

import torch
from torch import nn

from . import _lib, ops


def _grid_from_obs_dim(obs_dim, state_dim, rgb_dim):
    v = obs_dim - state_dim - rgb_dim
    g = round(v ** (1.0 / 3.0))
    if g ** 3 != v:
        raise ValueError(f"observation length {obs_dim} is not state({state_dim}) + G^3 + rgb({rgb_dim})")
    return g


class _EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, obs, enc, *params):
        feats = enc._run_forward(obs, need_bwd=True)     # grad mode is off inside Function.forward: say so explicitly
        ctx.enc, ctx.obs = enc, obs
        ctx.ws = enc._ws                     # the activations of THIS forward live here until backward
        enc._busy.add(ctx.ws.data_ptr())
        ctx.batch = obs.shape[0]
        ctx.training = enc.training
        ctx.save_for_backward(feats)
        return feats

    @staticmethod
    def backward(ctx, dfeat):
        enc = ctx.enc
        (feats,) = ctx.saved_tensors
        grads = enc._run_backward(ctx.obs, feats, dfeat.contiguous(), ctx.batch, ctx.training, ctx.ws)
        enc._busy.discard(ctx.ws.data_ptr())
        return (None, None) + tuple(grads)


class Hybrid_Encoder(nn.Module):
    RGB_DIM = 2 * 64 * 64          # k frames of 64x64 (env_train_gennbv.py:195-197), appended after the grid

    def __init__(self, observation_space, encoder_param=None, net_param=None, visual_input_shape=None,
                 state_input_shape=None, grid_size=None, semantic_branch=False):
        assert encoder_param is not None, "Need parameters !"
        assert net_param is not None, "Need parameters !"
        assert isinstance(visual_input_shape, (list, tuple)), "Use tuple or list"
        assert isinstance(state_input_shape, (list, tuple)), "Use tuple or list"
        super().__init__()
        self.image_channel = visual_input_shape[0]
        self.image_shape = visual_input_shape[1:]
        self.state_input_shape = state_input_shape
        feature_dim = net_param["append_hidden_shapes"][-1]
        net_param["append_hidden_shapes"].pop()
        self._observation_space = observation_space
        self._features_dim = feature_dim
        if feature_dim != 256:
            raise ValueError("the released architecture has features_dim = 256 (hybrid_encoder.py:52)")
        self.state_dim = int(state_input_shape[0])
        obs_dim = int(observation_space.shape[0])
        self.grid_size = int(grid_size) if grid_size is not None else _grid_from_obs_dim(obs_dim, self.state_dim, self.RGB_DIM)
        G = self.grid_size
        g1 = (G - 3) // 2 + 1
        g2 = (g1 - 3) // 2 + 1
        self.flat2 = 16 * g2 ** 3
        # parameter containers, reference construction order (hybrid_encoder.py:31-54)
        self.naive_encoder_grid = nn.Sequential(
            nn.Conv3d(1, 16, kernel_size=3, stride=2, padding=0), nn.BatchNorm3d(16), nn.ReLU(inplace=True),
            nn.Conv3d(16, 16, kernel_size=3, stride=2, padding=0), nn.BatchNorm3d(16), nn.ReLU(inplace=True))
        self.output_layer_grid = nn.Sequential(nn.Linear(self.flat2, 256), nn.ReLU(inplace=True))
        self.naive_encoder_action = nn.Sequential(nn.Linear(4 * self.state_dim, 256), nn.ReLU(inplace=True),
                                                  nn.Linear(256, 256), nn.ReLU(inplace=True))
        self.semantic_branch = bool(semantic_branch)
        self.output_layer = nn.Sequential(nn.Linear(768 if self.semantic_branch else 512, 256), nn.ReLU(inplace=True))
        if self.semantic_branch:             # created last: the first 16 tensors keep the reference's RNG stream and keys
            if obs_dim != self.state_dim + G ** 3 + self.RGB_DIM:
                raise ValueError("semantic_branch needs the k*64*64 frame columns behind the grid in the observation")
            self.naive_encoder_rgb = nn.Sequential(nn.Conv2d(2, 16, kernel_size=3, stride=2, padding=0), nn.ReLU(inplace=True),
                                                   nn.Conv2d(16, 16, kernel_size=3, stride=2, padding=0), nn.ReLU(inplace=True))
            self.output_layer_rgb = nn.Sequential(nn.Linear(16 * 15 * 15, 256), nn.ReLU(inplace=True))
        self._ws_cache = {}
        self._ws = None
        self._busy = set()                   # workspaces holding activations of a forward whose backward is pending

    @property
    def features_dim(self):
        return self._features_dim

    # order of the tensors handed to autograd (and returned by gnbv_encoder_backward)
    def _param_list(self):
        g, a = self.naive_encoder_grid, self.naive_encoder_action
        base = [g[0].weight, g[0].bias, g[1].weight, g[1].bias, g[3].weight, g[3].bias, g[4].weight, g[4].bias,
                self.output_layer_grid[0].weight, self.output_layer_grid[0].bias, a[0].weight, a[0].bias,
                a[2].weight, a[2].bias, self.output_layer[0].weight, self.output_layer[0].bias]
        if self.semantic_branch:
            r = self.naive_encoder_rgb
            base += [r[0].weight, r[0].bias, r[2].weight, r[2].bias, self.output_layer_rgb[0].weight, self.output_layer_rgb[0].bias]
        return base

    def _c_params(self):
        g, a = self.naive_encoder_grid, self.naive_encoder_action
        t = dict(conv1_w=g[0].weight, conv1_b=g[0].bias, bn1_w=g[1].weight, bn1_b=g[1].bias, bn1_rm=g[1].running_mean,
                 bn1_rv=g[1].running_var, bn1_nbt=g[1].num_batches_tracked, conv2_w=g[3].weight, conv2_b=g[3].bias,
                 bn2_w=g[4].weight, bn2_b=g[4].bias, bn2_rm=g[4].running_mean, bn2_rv=g[4].running_var,
                 bn2_nbt=g[4].num_batches_tracked, grid_fc_w=self.output_layer_grid[0].weight,
                 grid_fc_b=self.output_layer_grid[0].bias, act_fc1_w=a[0].weight, act_fc1_b=a[0].bias,
                 act_fc2_w=a[2].weight, act_fc2_b=a[2].bias, out_fc_w=self.output_layer[0].weight,
                 out_fc_b=self.output_layer[0].bias)
        if self.semantic_branch:
            r = self.naive_encoder_rgb
            t.update(rgb_conv1_w=r[0].weight, rgb_conv1_b=r[0].bias, rgb_conv2_w=r[2].weight, rgb_conv2_b=r[2].bias,
                     rgb_fc_w=self.output_layer_rgb[0].weight, rgb_fc_b=self.output_layer_rgb[0].bias)
        p = _lib.EncoderParams()
        for k, v in t.items():
            if not v.is_cuda or not v.is_contiguous():
                raise RuntimeError(f"Hybrid_Encoder parameter {k} must be a contiguous CUDA tensor (no CPU path)")
            want = torch.int64 if k.endswith("nbt") else torch.float32
            if v.dtype != want:
                raise RuntimeError(f"Hybrid_Encoder parameter {k}: dtype {v.dtype}, expected {want}")
            setattr(p, k, v.data_ptr())
        return p

    def _workspace(self, batch, device, with_backward):
        key = (batch, str(device), bool(with_backward))
        cached = self._ws_cache.get(key)
        if cached is None or cached.data_ptr() in self._busy:
            n = _lib.lib().gnbv_encoder_workspace_bytes(batch, self.grid_size, self.state_dim, int(with_backward))
            if n == 0:
                raise RuntimeError("gnbv_encoder_workspace_bytes rejected the sizes")
            fresh = torch.empty(n, dtype=torch.uint8, device=device)
            if cached is None:
                self._ws_cache[key] = fresh
            cached = fresh                   # a pending backward owns the cached one: use a private buffer
        self._ws = cached
        return self._ws

    def _check_obs(self, obs):
        if not obs.is_cuda or obs.dtype != torch.float32 or obs.dim() != 2:
            raise RuntimeError("Hybrid_Encoder.forward: expected a float32 CUDA tensor [N, D] (no CPU path)")
        if obs.stride(1) != 1:
            obs = obs.contiguous()
        if obs.shape[1] < self.state_dim + self.grid_size ** 3 + (self.RGB_DIM if self.semantic_branch else 0):
            raise RuntimeError("observation row shorter than state + grid (+ frames)")
        return obs

    def _run_forward(self, obs, need_bwd=False, row_index=None, training=None, feats=None):
        """obs [rows, D]; with row_index ([B] i64 on the device) sample b reads obs[row_index[b]] (no gather copy)."""
        B = obs.shape[0] if row_index is None else row_index.shape[0]
        training = self.training if training is None else training
        ws = self._workspace(B, obs.device, need_bwd)
        feats = torch.empty(B, 256, device=obs.device) if feats is None else feats
        p = self._c_params()
        rc = _lib.lib().gnbv_encoder_forward(ctypes.byref(p), obs.data_ptr(), obs.stride(0),
                                             None if row_index is None else row_index.data_ptr(), B, self.grid_size,
                                             self.state_dim, int(training), feats.data_ptr(), ws.data_ptr(),
                                             ws.numel(), ops._stream())
        _lib.check(rc, "gnbv_encoder_forward")
        return feats

    def _run_backward(self, obs, feats, dfeat, batch, ctx_training=True, ws=None, row_index=None, grads=None, phases=3):
        grads = [torch.empty_like(p) for p in self._param_list()] if grads is None else grads
        gp = _lib.EncoderGrads()
        for name, g in zip(_lib.EncoderGrads.FIELDS, grads):
            setattr(gp, name, g.data_ptr())
        p = self._c_params()
        ws = self._ws if ws is None else ws
        rc = _lib.lib().gnbv_encoder_backward_phase(ctypes.byref(p), obs.data_ptr(), obs.stride(0),
                                                    None if row_index is None else row_index.data_ptr(), batch, self.grid_size,
                                                    self.state_dim, int(ctx_training), feats.data_ptr(), dfeat.data_ptr(),
                                                    ctypes.byref(gp), ws.data_ptr(), ws.numel(), int(phases), ops._stream())
        _lib.check(rc, "gnbv_encoder_backward_phase")
        return grads

    def saved_activation(self, which, batch, ws=None):
        """Debug / test view of an activation the last forward left in its workspace (gnbv_encoder_workspace_view):
        which in {"y1", "stat1", "y2", "stat2", "act2"} -> a float32 tensor view."""
        ids = {"y1": 0, "stat1": 1, "y2": 2, "stat2": 3, "act2": 4}
        off, cnt = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(_lib.lib().gnbv_encoder_workspace_view(batch, self.grid_size, self.state_dim, ids[which], ctypes.byref(off),
                                                          ctypes.byref(cnt)), "gnbv_encoder_workspace_view")
        ws = self._ws if ws is None else ws
        return ws[off.value * 4:(off.value + cnt.value) * 4].view(torch.float32)

    def forward(self, observations):
        obs = self._check_obs(observations)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return _EncoderFn.apply(obs, self, *self._param_list())
        return self._run_forward(obs)

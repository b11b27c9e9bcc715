"""`Env_Train_GenNBV` -- drop-in for the reference task env (gennbv/env/env_train_gennbv.py) whose whole
step() runs as a handful of CUDA launches for all environments, with no host read-back.

Same public surface as the reference class: `reset() -> obs_dict`,
`step(actions[N,6] i64) -> (obs_dict, rew[N], done[N] bool, infos)`, attributes `num_envs, device,
observation_space, action_space, episode_length_buf, max_episode_length, extras, rew_buf, reset_buf,
time_out_buf, prob_grid, scanned_gt_grid, occ_grids_tri_cls, grid_gt, range_gt, voxel_size_gt,
num_valid_voxel_gt, poses, actions, ...`, `seed()`, `close()`.  Returned tensors are views of env-owned
buffers (the algorithm mutates `rew_buf` in place, on_policy_algorithm_grid_obs.py:208).

What differs by design:
  * Isaac Gym is replaced by an injectable `SensorSource` (gennbv_b200/sensors.py);
  * the observation lives pre-flattened in `obs_flat [N, D]` (state | grid | state_rgb, the order the
    reference wrapper concatenates, env_wrapper_gennbv_train.py:102-110); the dict entries are views of it
    and the tri-class grid is written there directly by the grid-update kernel;
  * pose / frame histories are [N,hist,6] / [N,k,64,64] tensors instead of deques of tensors;
  * episode statistics (infos["episode"]) stay on the device as 0-dim tensors.
"""
import math

import numpy as np
import torch

from . import _lib, ops, synth
from .config import Config_GenNBV_Train
from .sensors import SensorFrame, SensorSource
from .spaces import Box, Dict, MultiDiscrete


def _dptr(t):
    return None if t is None else t.data_ptr()


class _RatioHistory:
    """`reward_ratio_buf[-1]` / `[-2]` of the reference deque (env_train_gennbv.py:160-161): only the newest
    entry is ever consumed after the reward has been formed, so one tensor is kept."""

    def __init__(self, last):
        self._last = last

    def __getitem__(self, i):
        if i in (-1,):
            return self._last
        raise IndexError("only reward_ratio_buf[-1] is retained by gennbv_b200")


class Env_Train_GenNBV:
    # What `task_registry.make_env` (legged_gym/utils/task_registry.py:98-104) cannot pass: it calls
    # `task_class(cfg=, sim_params=, physics_engine=, sim_device=, headless=)` only.  A deployment installs the two factories once
    # (e.g. `Env_Train_GenNBV.sensor_factory = lambda env: IsaacGymSensor(...)`); explicit keyword arguments win.
    sensor_factory = None          # callable(env) -> SensorSource, called with cfg / num_envs / device already set
    grid_gt_loader = None          # callable(env) -> [num_scene, G, G, G, 4] tensor (reference: torch.load of the Houses3K GT grid)

    def __init__(self, cfg=None, sim_params=None, physics_engine=None, sim_device="cuda:0", headless=True, *,
                 sensor: SensorSource = None, grid_gt: torch.Tensor = None, num_envs: int = None):
        """cfg: Config_GenNBV_Train-like object; grid_gt: the [num_scene, G, G, G, 4] tensor the reference loads from
        data_gennbv/train/gt/train_houses3k_grid_gt.pt (env_train_gennbv.py:61-64); sensor: frame source.
        The buffers may be allocated on a non-CUDA device (construction only, e.g. to exercise an entry script's wiring):
        every kernel call checks its tensors and raises -- there is no CPU compute path."""
        _lib.lib()
        self.cfg = cfg if cfg is not None else Config_GenNBV_Train()
        cfg = self.cfg
        self.sim_params, self.physics_engine = sim_params, physics_engine
        self.device = torch.device(sim_device)
        self.headless = headless
        self.num_envs = int(num_envs if num_envs is not None else cfg.env.num_envs)
        if grid_gt is None and type(self).grid_gt_loader is not None:
            grid_gt = type(self).grid_gt_loader(self)
        if sensor is None and type(self).sensor_factory is not None:
            sensor = type(self).sensor_factory(self)
        self.sensor = sensor
        if sensor is None or grid_gt is None:
            raise ValueError("Env_Train_GenNBV needs a sensor source and the GT grid tensor (keyword arguments, or the "
                             "class-level sensor_factory / grid_gt_loader hooks)")
        N, dev = self.num_envs, self.device
        self.dt = cfg.control.decimation * float(np.float32(cfg.sim.dt))         # drone_robot.py:874-875
        self.max_episode_length = cfg.max_episode_length                          # env_train_base.py:132
        self.max_episode_length_s = cfg.env.episode_length_s                      # drone_robot.py:881
        self._load_gt(grid_gt)
        G = self.grid_size
        # ---- BaseTask buffers (base_task.py:73-91)
        self.rew_buf = torch.zeros(N, device=dev)
        self._reset_u8 = torch.ones(N, dtype=torch.uint8, device=dev)
        self._time_out_u8 = torch.zeros(N, dtype=torch.uint8, device=dev)
        self._time_outs_extra_u8 = torch.zeros(N, dtype=torch.uint8, device=dev)
        self.episode_length_buf = torch.zeros(N, dtype=torch.long, device=dev)
        self.extras = {}
        # ---- _init_buffers (env_train_gennbv.py:123-202)
        nz = cfg.normalization
        self.buffer_size = cfg.visual_input.stack
        self.action_size = len(nz.init_action)
        self._init_action = torch.tensor(nz.init_action, dtype=torch.long, device=dev)
        self.actions = self._init_action.repeat(N, 1).contiguous()
        self.action_unit = torch.tensor(nz.action_unit, device=dev)
        self.action_low_world = torch.tensor(nz.clip_pose_low, device=dev)
        self.clip_pose_idx_low = torch.tensor(nz.clip_pose_idx_low, dtype=torch.int64, device=dev)
        self.clip_pose_idx_up = torch.tensor(nz.clip_pose_idx_up, dtype=torch.int64, device=dev)
        self._init_pose = torch.tensor(nz.init_pose_buf, dtype=torch.float32, device=dev)
        self.poses = torch.zeros(N, self.action_size, device=dev)
        self.pose_hist = self._init_pose.repeat(N, self.buffer_size, 1).contiguous()     # oldest first
        self.k, self.rgb_h, self.rgb_w = 2, 64, 64
        self.rgb_hist = torch.zeros(N, self.k, self.rgb_h, self.rgb_w, device=dev)
        self.ratio_threshold_term = 0.99                   # check_termination (env_train_gennbv.py:454-457)
        self._accumulate_reset = False                     # `reset_buf = ...` here, `|=` in the eval env
        self._ratio = torch.zeros(N, device=dev)
        self.reward_ratio_buf = _RatioHistory(self._ratio)
        self.cur_reward_sum = torch.zeros(N, device=dev)
        self.cur_episode_length = torch.zeros(N, device=dev)
        self.H, self.W = sensor.height, sensor.width
        # inverse intrinsics: computed on the host like the reference does on its device (env_train_gennbv.py:168-169)
        self.inv_intri = torch.linalg.inv(synth.camera_intrinsics(self.H, self.W, cfg.visual_input.horizontal_fov)) \
            .to(torch.float32).contiguous().to(dev)
        self.blender2opencv = torch.tensor(synth.BLENDER2OPENCV, dtype=torch.float32)
        num_cols = int(math.floor(math.sqrt(N)))                                   # drone_robot.py:843-872
        num_rows = int(math.ceil(N / num_cols))
        xx, yy = torch.meshgrid(torch.arange(num_rows), torch.arange(num_cols), indexing="ij")
        self.env_origins = torch.zeros(N, 3)
        self.env_origins[:, 0] = cfg.env.env_spacing * xx.flatten()[:N]
        self.env_origins[:, 1] = cfg.env.env_spacing * yy.flatten()[:N]
        # ---- grids + flat observation
        self.scanned_gt_grid = torch.zeros(N, G, G, G, device=dev)
        self.prob_grid = torch.zeros(N, G, G, G, device=dev)
        V = G ** 3
        self._state_dim = self.buffer_size * self.action_size
        self._rgb_dim = self.k * self.rgb_h * self.rgb_w
        self.obs_dim = self._state_dim + V + self._rgb_dim
        # two observation / done buffers used alternately: the reference returns fresh tensors every step and SB3
        # keeps the previous step's obs and dones alive across the next env.step() (on_policy_algorithm_grid_obs.py:209-211)
        self._obs_pp = [torch.zeros(N, self.obs_dim, device=dev) for _ in range(2)]
        self._dones_pp = [torch.zeros(N, dtype=torch.uint8, device=dev) for _ in range(2)]
        self._pp = 0
        self._obs_target = None
        self.obs_flat = self._obs_pp[0]
        self._dones_u8 = self._dones_pp[0]
        self._cov_sum = torch.zeros(N, device=dev)
        self._sparse_update = (G ** 3) % 4 == 0 and self.obs_dim % 4 == 0 and self._state_dim % 4 == 0
        self._num_targets = torch.zeros(N, dtype=torch.int32, device=dev)
        self._workspace = ops.voxelize_workspace(N, G, dev)
        # ---- rewards (drone_robot.py:660-691: scale *= dt, zero scales dropped)
        sc = cfg.rewards.scales
        self.reward_scales = {k: getattr(sc, k) * self.dt for k in ("surface_coverage", "short_path", "termination")
                              if getattr(sc, k, 0) != 0}
        self.episode_sums_buf = torch.zeros(3, N, device=dev)
        self.episode_sums = {"surface_coverage": self.episode_sums_buf[0], "short_path": self.episode_sums_buf[1],
                             "termination": self.episode_sums_buf[2]}
        self._stats = torch.zeros(int(_lib.lib().gnbv_episode_stats_doubles()), dtype=torch.float64, device=dev)
        self.profile_events = None
        self.update_observation_space()

    # ------------------------------------------------------------------ GT (env_train_gennbv.py:56-96)
    def _load_gt(self, grid_gt):
        N, dev = self.num_envs, self.device
        grid_gt = grid_gt.float()
        vs, nvalid, rg = synth.gt_metadata(grid_gt)
        self.num_scene = grid_gt.shape[0]
        self.grid_size = grid_gt.shape[1]
        assert grid_gt.shape[1] == grid_gt.shape[2] == grid_gt.shape[3]
        self.env_to_scene = (torch.arange(N) % self.num_scene)
        idx = self.env_to_scene
        self.grid_gt = grid_gt[..., 3][idx].contiguous().to(dev)
        self.voxel_size_gt = vs[idx].contiguous().to(dev)
        self.num_valid_voxel_gt = nvalid[idx].contiguous().to(dev)
        self.range_gt = rg[idx].contiguous().to(dev)
        self.env_to_scene = idx.to(dev)

    def update_observation_space(self):
        """env_train_gennbv.py:459-492."""
        size = (self.clip_pose_idx_up - self.clip_pose_idx_low + 1).cpu().numpy()
        self.action_space = MultiDiscrete(nvec=size)
        rg = self.range_gt.cpu()
        up = [rg[:, 0].max().item(), rg[:, 2].max().item(), rg[:, 4].max().item(), 0, 1 / 2 * np.pi, 2 * np.pi]
        low = [rg[:, 1].min().item(), rg[:, 3].min().item(), rg[:, 5].min().item(), 0, -1 / 2 * np.pi, 0]
        G = self.grid_size
        self.observation_space = Dict({
            "state": Box(low=np.tile(low, self.buffer_size).astype(np.float32),
                         high=np.tile(up, self.buffer_size).astype(np.float32),
                         shape=(self.buffer_size * self.action_size,), dtype=np.int64),
            "state_rgb": Box(low=0, high=255, shape=(self.k * self.rgb_h * self.rgb_w,), dtype=np.int64),
            "grid": Box(low=-np.inf, high=np.inf, shape=(G, G, G), dtype=np.float32),
        })

    # ------------------------------------------------------------------ views with the reference's names
    @property
    def reset_buf(self):
        return self._reset_u8.view(torch.bool)

    @property
    def time_out_buf(self):
        return self._time_out_u8.view(torch.bool)

    @property
    def occ_grids_tri_cls(self):
        G = self.grid_size
        return self.obs_flat[:, self._state_dim:self._state_dim + G ** 3].view(self.num_envs, G, G, G)

    def _obs_dict(self):
        N, s, V = self.num_envs, self._state_dim, self.grid_size ** 3
        return {"state": self.obs_flat[:, :s].view(N, self.buffer_size, self.action_size),
                "state_rgb": self.obs_flat[:, s + V:].view(N, self.k, self.rgb_h, self.rgb_w),
                "grid": self.occ_grids_tri_cls}

    def get_pose_from_discrete_action(self, action):
        return action * self.action_unit + self.action_low_world               # env_train_base.py:665-667

    def seed(self, seed):
        torch.manual_seed(seed)
        np.random.seed(seed)

    def close(self):
        pass

    def render(self, *a, **k):
        pass

    # ------------------------------------------------------------------ step / reset
    def step(self, actions):
        """env_train_gennbv.py:246-264."""
        L, s = _lib.lib(), ops._stream()
        if actions.dtype != torch.int64 or not actions.is_cuda:
            raise RuntimeError("step(actions): expected an int64 CUDA tensor [N,6]")
        actions = actions.contiguous()
        _lib.check(L.gnbv_actions_to_poses(actions.data_ptr(), self.episode_length_buf.data_ptr(),
                                           self.clip_pose_idx_low.data_ptr(), self.clip_pose_idx_up.data_ptr(),
                                           self._init_action.data_ptr(), self.action_unit.data_ptr(),
                                           self.action_low_world.data_ptr(), self.actions.data_ptr(), self.poses.data_ptr(),
                                           self.num_envs, self.action_size, s), "gnbv_actions_to_poses")
        return self.post_physics_step()

    def reset(self):
        """env_train_gennbv.py:229-244: reset every env, then one full observation pass from the initial pose."""
        self._reset_u8.fill_(1)
        self._time_outs_extra_u8.copy_(self._time_out_u8)        # reset_idx binds extras["time_outs"] (:435-436)
        self._stats[204:207] = self.episode_sums_buf.double().mean(dim=1) / self.max_episode_length_s     # :424-428
        self._reset_flagged(clear=False)
        self.actions.copy_(torch.clip(self.actions, self.clip_pose_idx_low, self.clip_pose_idx_up))
        self.poses.copy_(self.get_pose_from_discrete_action(self.actions))
        return self.post_physics_step(if_reset=True)

    def _reset_flagged(self, clear):
        L = _lib.lib()
        if self._sparse_update:
            self._cov_sum.masked_fill_(self._reset_u8.view(torch.bool), 0.0)       # carried coverage sum of the envs being reset
        _lib.check(L.gnbv_reset_envs(self._reset_u8.data_ptr(), self.prob_grid.data_ptr(), self.scanned_gt_grid.data_ptr(),
                                     self.pose_hist.data_ptr(), self.rgb_hist.data_ptr(), self._ratio.data_ptr(),
                                     self.actions.data_ptr(), self.episode_length_buf.data_ptr(),
                                     self.episode_sums_buf.data_ptr(), self._init_pose.data_ptr(),
                                     self._init_action.data_ptr(), self.num_envs, self.grid_size, self.buffer_size,
                                     self.action_size, self.k, self.rgb_h, self.rgb_w, int(clear), ops._stream()),
                   "gnbv_reset_envs")

    def _c2w(self, frame: SensorFrame):
        if frame.c2w is not None:
            return frame.c2w
        # the reference's own lines (env_train_gennbv.py:512-514) on the host array Isaac Gym returns, then one
        # asynchronous H2D copy of N x 64 bytes
        ext = torch.from_numpy(frame.view_matrix)
        c2w = torch.linalg.inv(ext.transpose(-2, -1)) @ self.blender2opencv.unsqueeze(0)
        c2w[:, :3, 3] -= self.env_origins
        return c2w.contiguous().to(self.device, non_blocking=True)

    def bind_next_observation(self, target):
        """The next step()/reset() writes its flattened observation [N, D] straight into `target` (e.g. the rollout
        buffer's slot for the next transition) instead of the env's own ping-pong buffer; the returned
        observation is then a view of `target`.  One-shot; every column of the row is rewritten by the step."""
        if not (isinstance(target, torch.Tensor) and target.is_cuda and target.dtype == torch.float32 and target.is_contiguous()
                and tuple(target.shape) == (self.num_envs, self.obs_dim)):
            raise RuntimeError(f"bind_next_observation: expected a contiguous float32 CUDA tensor [{self.num_envs}, {self.obs_dim}]")
        self._obs_target = target

    # hooks of the eval env (gennbv_b200/env_eval.py); no-ops here
    def _after_occ_grid_update(self, frame, c2w):
        pass

    def _before_reset_idx(self):
        pass

    def post_physics_step(self, if_reset=False):
        """env_train_gennbv.py:328-375 (post_physics_step + get_step_return) as 6 launches."""
        L, s, N, G = _lib.lib(), ops._stream(), self.num_envs, self.grid_size
        self._pp ^= 1
        self.obs_flat, self._dones_u8 = self._obs_pp[self._pp], self._dones_pp[self._pp]
        if self._obs_target is not None:                 # one-shot: this step's observation is born in the caller's storage
            self.obs_flat, self._obs_target = self._obs_target, None
        frame = self.sensor.render(self.poses)
        self._frame = frame
        c2w = self._c2w(frame)
        rgba = frame.rgba
        if rgba is not None and (rgba.dtype != torch.uint8 or not rgba.is_contiguous()):
            raise RuntimeError("sensor rgba must be a contiguous uint8 [N,H,W,4] tensor")
        # update_obs_buf + state / state_rgb columns of the observation
        _lib.check(L.gnbv_obs_update(_dptr(rgba), self.poses.data_ptr(), self.pose_hist.data_ptr(), self.rgb_hist.data_ptr(),
                                     self.obs_flat.data_ptr(), self.obs_dim, 0, self._state_dim + G ** 3, N, self.H, self.W,
                                     self.buffer_size, self.action_size, self.k, self.rgb_h, self.rgb_w, s),
                   "gnbv_obs_update")
        # update_occ_grid: tri-class grid lands in the grid columns of obs_flat
        self._xyz = self.poses[:, :3].contiguous()
        ev = self.profile_events                      # optional (start, mid, end) CUDA events around the two voxelize phases
        if ev is not None:
            ev[0].record()
        ops.scan_raycast(frame.depth, frame.seg, self.inv_intri, c2w, self.range_gt, self.voxel_size_gt, self._xyz, G,
                         self._workspace, self._num_targets, raw_depth=True)
        if ev is not None:
            ev[1].record()
        # sparse update: groups of voxels no ray touched are not rewritten, the coverage sum is carried across steps
        ops.grid_update(self.grid_gt, self.prob_grid, self.scanned_gt_grid, self.obs_flat.view(-1)[self._state_dim:],
                        self._cov_sum, self._workspace, tri_row_stride=self.obs_dim, sparse=self._sparse_update)
        if ev is not None:
            ev[2].record()
        self._after_occ_grid_update(frame, c2w)
        # compute_reward / check_termination / episode statistics
        rs = self.reward_scales
        _lib.check(L.gnbv_reward_termination(
            self._cov_sum.data_ptr(), self.num_valid_voxel_gt.data_ptr(), self._ratio.data_ptr(),
            self.episode_length_buf.data_ptr(), _dptr(frame.contact) if self.cfg.termination.collision else None,
            self.rew_buf.data_ptr(), self._reset_u8.data_ptr(), self._time_out_u8.data_ptr(), self._dones_u8.data_ptr(),
            self.episode_sums_buf.data_ptr(), self.cur_reward_sum.data_ptr(), self.cur_episode_length.data_ptr(),
            self._stats.data_ptr(), self._time_outs_extra_u8.data_ptr(),
            float(rs.get("surface_coverage", 0.0)), float(rs.get("short_path", 0.0)), float(rs.get("termination", 0.0)),
            int("termination" in rs), int(bool(self.cfg.rewards.only_positive_rewards)),
            int(bool(self.cfg.termination.max_step_done)), int(self.max_episode_length), float(self.max_episode_length_s),
            float(self.ratio_threshold_term), int(self._accumulate_reset), N, s), "gnbv_reward_termination")
        obs = self._obs_dict()
        self._before_reset_idx()
        # reset_idx for the done envs (the returned observation is the pre-reset one, as in the reference)
        self._reset_flagged(clear=True)
        st = self._stats
        self.extras["episode"] = {"rew_surface_coverage": st[204], "rew_short_path": st[205], "rew_termination": st[206],
                                  "episode_reward": st[202], "episode_length": st[203]}
        self.extras["time_outs"] = self._time_outs_extra_u8.view(torch.bool)
        if if_reset:
            return obs
        return obs, self.rew_buf, self._dones_u8.view(torch.bool), self.extras

"""TEST HARNESS ONLY -- runs the reference's *own* `Env_Train_GenNBV` methods on CPU.

The reference env needs Isaac Gym to be constructed.  Here an instance is created with
`object.__new__` (no `__init__`), its attribute surface is filled the way
`_init_buffers` / `BaseTask.__init__` / `_prepare_reward_function` would fill it
(env_train_gennbv.py:98-202, base_task.py:42-105, drone_robot.py:660-691), Isaac Gym is a
MagicMock whose `render_all_camera_sensors` paints synthetic images
(gennbv_b200.synth) into the per-env camera tensors, and the PyCUDA Bresenham launcher --
which cannot run on CPU -- is swapped for the C restatement in oracle/gennbv_oracle.c.
Everything else that executes (`step`, `post_physics_step`, `get_step_return`,
`post_process_camera_tensor`, `update_occ_grid`, `back_projection_fg`,
`scanned_pts_to_idx_3D`, `pose_coord_to_idx_3D`, `grid_occupancy_tri_cls`,
`compute_reward`, `check_termination`, `reset_idx`, ...) is the reference's unmodified code.

Only usable where /root/reference exists (the build container).
"""
import os
import sys
from collections import deque
from unittest.mock import MagicMock, patch

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

import oracle as c_oracle  # noqa: E402  (oracle/oracle.py)
import ref_loader  # noqa: E402
from gennbv_b200 import synth  # noqa: E402  (input synthesis only)

SIM_DT = float(np.float32(0.005))   # gymapi.SimParams.dt is a C float (drone_robot.py:875 comment)
DECIMATION = 4


def _bresenham_cpu(pts_source, pts_target, map_size):
    """Drop-in for gennbv.utils.bresenham3D_pycuda (utils.py:24-227) backed by the C restatement."""
    if isinstance(map_size, list):
        map_size = map_size[0]
    out = c_oracle.bresenham3d(pts_source.int().numpy().astype(np.int64),
                               pts_target.int().numpy().astype(np.int64), int(map_size))
    return torch.from_numpy(out).to(torch.long)


def chamfer_distance_restated(x, y):
    """Stand-in for pytorch3d.loss.chamfer_distance (third party, "0.7.8 works" per the reference README, not vendored,
    absent here) with its defaults as called at env_eval_gennbv.py:258: squared L2, point_reduction="mean",
    batch_reduction="mean", both directions summed; float64 brute force.  Returns (loss, None) like pytorch3d."""
    d = torch.cdist(x.double(), y.double()) ** 2                 # [B,P1,P2]
    loss = d.min(2).values.mean(1) + d.min(1).values.mean(1)     # [B]
    return loss.mean().float(), None


def make_reference_env(num_envs, H, W, scenes, buffer_size=100, max_episode_length=100, with_rgb=True, eval_env=False,
                       pc_gt=None):
    """-> (env, ref) where env is a CPU instance of the reference's Env_Train_GenNBV (or, with eval_env=True, of its
    Env_Eval_GenNBV with `pc_gt` = list of per-env GT clouds and pytorch3d's chamfer replaced by the restatement above)."""
    ref = ref_loader.load_reference()
    ref.env_train.bresenham3D_pycuda = _bresenham_cpu
    if eval_env:
        env_eval_mod = __import__("gennbv.env.env_eval_gennbv", fromlist=["x"])
        env_eval_mod.bresenham3D_pycuda = _bresenham_cpu
        env_eval_mod.chamfer_distance = chamfer_distance_restated
        Env = env_eval_mod.Env_Eval_GenNBV
        from gennbv.env.config_gennbv_eval import Config_GenNBV_Eval as Config
    else:
        Env = ref.env_train.Env_Train_GenNBV
        from gennbv.env.config_gennbv_train import Config_GenNBV_Train as Config

    cfg = Config()
    cfg.max_episode_length = max_episode_length
    cfg.visual_input.stack = buffer_size
    cfg.visual_input.camera_height, cfg.visual_input.camera_width = H, W
    cfg.env.num_envs = num_envs
    if not eval_env:
        cfg.rewards.only_positive_rewards = False     # train_gennbv.py:101-106 overrides the config
    cfg.return_visual_observation = True
    if cfg.terrain.mesh_type not in ["heightfield", "trimesh"]:     # drone_robot.py:879-880
        cfg.terrain.curriculum = False

    env = object.__new__(Env)
    env.cfg = cfg
    env.device = "cpu"
    env.num_envs = num_envs
    env.num_scene = scenes.num_scenes
    env.gym, env.sim, env.viewer = MagicMock(), MagicMock(), None
    env.enable_viewer_sync, env.debug_viz, env.headless = True, False, True
    env.dt = DECIMATION * SIM_DT
    env.max_episode_length = cfg.max_episode_length          # env_train_base.py:132
    env.max_episode_length_s = cfg.env.episode_length_s      # drone_robot.py:881
    # BaseTask.__init__ buffers (base_task.py:73-91)
    env.rew_buf = torch.zeros(num_envs)
    env.reset_buf = torch.ones(num_envs, dtype=torch.long)
    env.episode_length_buf = torch.zeros(num_envs, dtype=torch.long)
    env.time_out_buf = torch.zeros(num_envs, dtype=torch.bool)
    env.extras = {}
    # env origins grid (drone_robot.py:843-872)
    num_cols = int(np.floor(np.sqrt(num_envs)))
    num_rows = int(np.ceil(num_envs / num_cols))
    xx, yy = torch.meshgrid(torch.arange(num_rows), torch.arange(num_cols), indexing="ij")
    env.env_origins = torch.zeros(num_envs, 3)
    env.env_origins[:, 0] = cfg.env.env_spacing * xx.flatten()[:num_envs]
    env.env_origins[:, 1] = cfg.env.env_spacing * yy.flatten()[:num_envs]
    # GT: run the reference's _init_load_all on the synthetic file content
    def fake_load(path, *a, **k):
        name = os.path.basename(str(path))
        if name.endswith("_pc.pt"):                              # BAT12_SETA_HOUSE{env+1}_pc.pt (env_eval_gennbv.py:95-101)
            return pc_gt[int("".join(ch for ch in name.split("HOUSE")[1] if ch.isdigit())) - 1].clone()
        return scenes.grid_gt.clone()

    with patch.object(torch, "load", fake_load):
        env._init_load_all()
    if eval_env:
        env.ratios_accuracy = dict()                             # Env_Eval_GenNBV._init_buffers (:103-105)
    # the part of _init_buffers that does not touch Isaac Gym (env_train_gennbv.py:123-202)
    env.contact_forces = torch.zeros(num_envs, 6, 3)
    env.termination_contact_indices = torch.tensor([0, 2, 3, 4, 5])
    env.penalised_contact_indices = torch.tensor([0, 2, 3, 4, 5])
    env.rewbuffer, env.lenbuffer = deque(maxlen=100), deque(maxlen=100)
    env.buffer_size = cfg.visual_input.stack
    env.cur_reward_sum = torch.zeros(num_envs)
    env.cur_episode_length = torch.zeros(num_envs)
    nz = cfg.normalization
    env.actions = torch.tensor(nz.init_action, dtype=torch.long).repeat(num_envs, 1)
    env.action_unit = torch.tensor(nz.action_unit)
    env.action_size = env.actions.shape[1]
    env.action_low_world = torch.tensor(nz.clip_pose_low)
    env.clip_pose_idx_low = torch.tensor(nz.clip_pose_idx_low, dtype=torch.int64)
    env.clip_pose_idx_up = torch.tensor(nz.clip_pose_idx_up, dtype=torch.int64)
    pose_buf = torch.tensor(nz.init_pose_buf, dtype=torch.float).repeat(num_envs, 1)
    env.pose_buf = deque(maxlen=env.buffer_size)
    env.pose_buf.extend(env.buffer_size * [pose_buf])
    env.ratio_threshold_term = 0.99
    env.reward_ratio_buf = deque(maxlen=max(env.buffer_size, 2))
    env.reward_ratio_buf.extend(max(env.buffer_size, 2) * [torch.zeros(num_envs)])
    env.collision_buf = torch.ones(num_envs, dtype=torch.long)
    env.blender2opencv = torch.FloatTensor(synth.BLENDER2OPENCV)
    env.inv_intri = torch.linalg.inv(env.get_camera_intrinsics()).to(torch.float32)
    xs = torch.linspace(0, W - 1, W, dtype=torch.float32)
    ys = torch.linspace(0, H - 1, H, dtype=torch.float32)
    ys, xs = torch.meshgrid(ys, xs, indexing="ij")
    ncp = torch.stack([xs, ys], dim=-1)
    env.norm_coord_pixel = torch.concat((ncp, torch.ones_like(ncp[..., :1])), dim=-1).view(-1, 3)
    G = env.grid_size
    env.scanned_gt_grid = torch.zeros(num_envs, G, G, G)
    env.prob_grid = torch.zeros(num_envs, G, G, G)
    env.occ_grids_tri_cls = torch.zeros(num_envs, G, G, G)
    env.k, env.rgb_h, env.rgb_w = 2, 64, 64
    env.rgb_buf = deque(maxlen=env.k)
    env.rgb_buf.extend(env.k * [torch.zeros((num_envs, 1, env.rgb_h, env.rgb_w))])
    env.pts_target_list = []
    # camera tensors Isaac Gym would own and refresh in place (env_train_gennbv.py:204-227)
    env.rgb_cam_tensors = [torch.zeros(H, W, 4, dtype=torch.uint8) for _ in range(num_envs)]
    env.depth_cam_tensors = [torch.zeros(H, W) for _ in range(num_envs)]
    env.seg_cam_tensors = [torch.zeros(H, W, dtype=torch.int32) for _ in range(num_envs)]
    env._view_matrix = np.zeros((num_envs, 4, 4), np.float32)
    env._last_c2w = None

    def paint(_sim):
        depth, seg, rgb, c2w = synth.render(scenes.params, env.poses, H, W,
                                            cfg.visual_input.horizontal_fov, with_rgb=with_rgb)
        for i in range(num_envs):
            env.depth_cam_tensors[i].copy_(depth[i])
            env.seg_cam_tensors[i].copy_(seg[i])
            if rgb is not None:
                env.rgb_cam_tensors[i].copy_(rgb[i])
        env._view_matrix = synth.c2w_to_isaac_view_matrix(c2w, env.env_origins).numpy()

    env.gym.render_all_camera_sensors.side_effect = paint
    env.get_camera_view_matrix = lambda: env._view_matrix
    env.render = lambda *a, **k: None
    env.set_state = lambda *a, **k: None
    env._reset_root_states = lambda *a, **k: None
    # reward plumbing (drone_robot.py:660-691, 875)
    from legged_gym.utils.helpers import class_to_dict
    env.reward_scales = class_to_dict(cfg.rewards.scales)
    env._prepare_reward_function()
    env.update_observation_space()
    return env, ref

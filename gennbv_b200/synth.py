"""Synthetic stand-in for the simulator side of GenNBV (Isaac Gym + Houses3K data).

The reference renders depth / segmentation / rgb with Isaac Gym camera sensors
(`gennbv/env/env_train_gennbv.py:346-354`, `env_train_base.py:513-534`) and loads
`data_gennbv/train/gt/train_houses3k_grid_gt.pt` ([num_scene, G, G, G, 4] = voxel
centre xyz + occupancy, `env_train_gennbv.py:56-96`).  Neither exists here, so
every BASELINE.json config is defined on synthetic inputs of the same shape and
dtype:

  * scenes  : analytic "houses" -- axis-aligned box with a gable roof, centred at
              the origin on z = 0; GT grid in the reference file layout;
  * sensors : pin-hole z-depth by exact ray / convex-polyhedron intersection,
              delivered in the *raw* Isaac Gym convention (negative depth, -inf for
              no hit; seg = 255 on the object, 0 elsewhere; RGBA uint8), plus the
              per-env view matrix in Isaac's transposed (row-vector) convention.

Everything is seeded `torch.Generator` arithmetic; the same code runs on CPU or
on the device.  This module is input synthesis, not part of the measured path.
"""
import math
from dataclasses import dataclass

import torch

# pose lattice of the reference: config_gennbv_train.py:62-69
CLIP_POSE_LOW = (-8.0, -8.0, 0.1, 0.0, -0.5 * 3.14159265359, 0.0)
CLIP_POSE_IDX_UP = (80, 80, 50, 0, 12, 12)
CLIP_POSE_IDX_LOW = (0, 0, 0, 0, 0, 0)
INIT_ACTION = (40, 40, 50, 0, 12, 0)
INIT_POSE_BUF = (0.0, 0.0, 10.1, 0.0, 90 / 180 * math.pi, 0.0)
ACTION_UNIT = (0.2, 0.2, 0.2, 0.0, 1 / 12 * math.pi, 1 / 6 * math.pi)

OBJECT_SEGMENTATION_ID = 255  # env_train_base.py:25
BLENDER2OPENCV = ((1.0, 0, 0, 0), (0, -1.0, 0, 0), (0, 0, -1.0, 0), (0, 0, 0, 1.0))


def camera_intrinsics(H: int, W: int, horizontal_fov_deg: float = 90.0) -> torch.Tensor:
    """K [3,3] float32, same arithmetic as `Env_Train_Base.get_camera_intrinsics`
    (env_train_base.py:787-803): fy is derived from FOV_y = FOV_x * H / W."""
    fov_x = horizontal_fov_deg / 180 * math.pi
    fov_y = fov_x * H / W
    fx = 0.5 * W / math.tan(0.5 * fov_x)
    fy = 0.5 * H / math.tan(0.5 * fov_y)
    return torch.tensor([[fx, 0, W / 2], [0, fy, H / 2], [0, 0, 1]]).float()


@dataclass
class HouseScenes:
    """S analytic houses + their GT grids in the reference file layout."""
    params: torch.Tensor      # [S,4] float64: lx, ly, wall height, roof height
    grid_gt: torch.Tensor     # [S,G,G,G,4] float32: centre x,y,z, occupancy {0,1}

    @property
    def num_scenes(self):
        return self.params.shape[0]


def _surface_samples(lx, ly, hw, hr, spacing):
    """Dense float64 samples on the visible faces (4 walls incl. gable ends, 2 roof slopes)."""
    def lin(a, b):
        n = max(int(math.ceil((b - a) / spacing)) + 1, 2)
        return torch.linspace(a, b, n, dtype=torch.float64)

    def roofz(y):
        return hw + hr * (1.0 - y.abs() / (ly / 2))

    pts = []
    xs, ys = lin(-lx / 2, lx / 2), lin(-ly / 2, ly / 2)
    zs_wall, zs_full = lin(0.0, hw), lin(0.0, hw + hr)
    # walls y = +-ly/2 (rectangles)
    X, Z = torch.meshgrid(xs, zs_wall, indexing="ij")
    for s in (-1.0, 1.0):
        pts.append(torch.stack([X, torch.full_like(X, s * ly / 2), Z], -1).reshape(-1, 3))
    # gable walls x = +-lx/2 (pentagons)
    Y, Z = torch.meshgrid(ys, zs_full, indexing="ij")
    keep = Z <= roofz(Y)
    for s in (-1.0, 1.0):
        pts.append(torch.stack([torch.full_like(Y, s * lx / 2), Y, Z], -1)[keep])
    # roof slopes
    X, Y = torch.meshgrid(xs, ys, indexing="ij")
    pts.append(torch.stack([X, Y, roofz(Y)], -1).reshape(-1, 3))
    return torch.cat(pts, 0)


def make_house_scenes(num_scenes: int, G: int, seed: int = 0) -> HouseScenes:
    """Footprint U[2,6] m, wall U[1.5,4] m, roof U[0.5,2] m; grid extent slightly larger
    than the bounding box by a non-dyadic factor so voxel faces are not representable
    depths."""
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(num_scenes, 4, generator=g, dtype=torch.float64)
    params = torch.stack([2 + 4 * u[:, 0], 2 + 4 * u[:, 1], 1.5 + 2.5 * u[:, 2], 0.5 + 1.5 * u[:, 3]], -1)
    grid = torch.zeros(num_scenes, G, G, G, 4, dtype=torch.float32)
    for s in range(num_scenes):
        lx, ly, hw, hr = (float(v) for v in params[s])
        xr, yr, zr = lx * 1.043, ly * 1.043, (hw + hr) * 1.021
        cx = torch.linspace(-xr / 2, xr / 2, G, dtype=torch.float64)
        cy = torch.linspace(-yr / 2, yr / 2, G, dtype=torch.float64)
        cz = torch.linspace(0.0, zr, G, dtype=torch.float64)
        CX, CY, CZ = torch.meshgrid(cx, cy, cz, indexing="ij")
        grid[s, ..., 0], grid[s, ..., 1], grid[s, ..., 2] = CX.float(), CY.float(), CZ.float()
        vs = torch.tensor([xr, yr, zr], dtype=torch.float64) / (G - 1)
        lo = torch.tensor([-xr / 2, -yr / 2, 0.0], dtype=torch.float64) - 0.5 * vs
        pts = _surface_samples(lx, ly, hw, hr, float(vs.min()) / 3.0)
        idx = torch.floor((pts - lo) / vs).long().clamp_(0, G - 1)
        occ = torch.zeros(G * G * G, dtype=torch.float32)
        occ[(idx[:, 0] * G + idx[:, 1]) * G + idx[:, 2]] = 1.0
        grid[s, ..., 3] = occ.view(G, G, G)
    return HouseScenes(params=params, grid_gt=grid)


def gt_point_clouds(params: torch.Tensor, num_envs: int, num_points: int, seed: int = 0):
    """Per-env GT surface clouds (list of [num_points,3] float32), the stand-in for the reference's
    data_gennbv/eval/gt/point_cloud/BAT12_SETA_HOUSE{e+1}_pc.pt files (env_eval_gennbv.py:93-101): a seeded random subset
    of dense samples on the visible faces of scene e % S."""
    g = torch.Generator().manual_seed(seed)
    S = params.shape[0]
    out = []
    for e in range(num_envs):
        lx, ly, hw, hr = (float(v) for v in params[e % S])
        area = 2 * lx * hw + 2 * ly * (hw + hr / 2) + 2 * lx * math.hypot(ly / 2, hr)
        dense = _surface_samples(lx, ly, hw, hr, max(math.sqrt(area / (4.0 * num_points)), 1e-3))
        while dense.shape[0] < num_points:
            dense = torch.cat([dense, dense + 1e-4], 0)
        sel = torch.randperm(dense.shape[0], generator=g)[:num_points]
        out.append(dense[sel].float().contiguous())
    return out


def gt_metadata(grid_gt: torch.Tensor):
    """voxel_size_gt [S,3], num_valid_voxel_gt [S], range_gt [S,6] -- the derivations of
    `Env_Train_GenNBV._init_load_all` (env_train_gennbv.py:66-80), same fp32 arithmetic."""
    voxel_size = torch.cat([grid_gt[:, 1, 0, 0, 0:1] - grid_gt[:, 0, 0, 0, 0:1],
                            grid_gt[:, 0, 1, 0, 1:2] - grid_gt[:, 0, 0, 0, 1:2],
                            grid_gt[:, 0, 0, 1, 2:3] - grid_gt[:, 0, 0, 0, 2:3]], dim=-1)
    num_valid = grid_gt[..., 3].sum(dim=(-1, -2, -3))
    xr = grid_gt[:, -1, 0, 0, 0:1] - grid_gt[:, 0, 0, 0, 0:1]
    yr = grid_gt[:, 0, -1, 0, 1:2] - grid_gt[:, 0, 0, 0, 1:2]
    zr = grid_gt[:, 0, 0, -1, 2:3] - grid_gt[:, 0, 0, 0, 2:3]
    range_gt = torch.cat([xr / 2, -xr / 2, yr / 2, -yr / 2, zr, torch.zeros_like(zr)], dim=-1)
    return voxel_size, num_valid, range_gt


def pose_from_action(action_idx: torch.Tensor) -> torch.Tensor:
    """idx * unit + low (env_train_base.py:665-667) in float32."""
    unit = torch.tensor(ACTION_UNIT, dtype=torch.float32, device=action_idx.device)
    low = torch.tensor(CLIP_POSE_LOW, dtype=torch.float32, device=action_idx.device)
    return action_idx * unit + low


def sample_lookat_actions(params: torch.Tensor, num_envs: int, generator: torch.Generator) -> torch.Tensor:
    """Lattice actions [N,6] int64 whose camera looks (up to lattice snapping) at the house
    of scene `env % S`, from 4.5-9 m away and outside its bounding box."""
    S = params.shape[0]
    out = torch.zeros(num_envs, 6, dtype=torch.int64)
    for e in range(num_envs):
        lx, ly, hw, hr = (float(v) for v in params[e % S])
        zc = 0.5 * (hw + hr)
        while True:
            ix = int(torch.randint(0, 81, (1,), generator=generator))
            iy = int(torch.randint(0, 81, (1,), generator=generator))
            iz = int(torch.randint(0, 51, (1,), generator=generator))
            x, y, z = ix * 0.2 - 8.0, iy * 0.2 - 8.0, iz * 0.2 + 0.1
            dist = math.sqrt(x * x + y * y + (z - zc) ** 2)
            inside_margin = abs(x) < lx / 2 + 0.6 and abs(y) < ly / 2 + 0.6 and z < hw + hr + 0.6
            if 4.5 <= dist <= 9.0 and not inside_margin:
                break
        yaw = math.atan2(-y, -x) % (2 * math.pi)
        pitch = math.atan2(z - zc, math.hypot(x, y))
        iyaw = int(round(yaw / (math.pi / 6))) % 12
        ipitch = min(max(int(round((pitch + math.pi / 2) / (math.pi / 12))), 0), 12)
        out[e] = torch.tensor([ix, iy, iz, 0, ipitch, iyaw])
    return out


def pose_to_c2w(poses: torch.Tensor) -> torch.Tensor:
    """[N,6] (x,y,z,roll,pitch,yaw) -> camera-to-world [N,4,4] float64, OpenCV camera axes
    (x right, y down, z forward).  Body frame: x forward, y left, z up; R = Rz(yaw) Ry(pitch)
    Rx(roll); positive pitch looks down (the reference's init pose, pitch = +90 deg at
    z = 10.1, looks straight down: config_gennbv_train.py:67)."""
    p = poses.double()
    r, pt, yw = p[:, 3], p[:, 4], p[:, 5]
    cr, sr, cp, sp, cy, sy = r.cos(), r.sin(), pt.cos(), pt.sin(), yw.cos(), yw.sin()
    R = torch.stack([
        torch.stack([cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr], -1),
        torch.stack([sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr], -1),
        torch.stack([-sp, cp * sr, cp * cr], -1)], 1)                          # body -> world
    cam_in_body = torch.tensor([[0.0, 0, 1], [-1, 0, 0], [0, -1, 0]], dtype=torch.float64, device=p.device)
    c2w = torch.eye(4, dtype=torch.float64, device=p.device).repeat(p.shape[0], 1, 1)
    c2w[:, :3, :3] = R @ cam_in_body
    c2w[:, :3, 3] = p[:, :3]
    return c2w


def c2w_to_isaac_view_matrix(c2w: torch.Tensor, env_origins: torch.Tensor) -> torch.Tensor:
    """What `gym.get_camera_view_matrix` would hand the reference: V such that
    inv(V^T) @ blender2opencv = c2w (+ env origin), i.e. the inverse of
    `back_projection_fg` lines 512-514.  float32 [N,4,4]."""
    b2o = torch.tensor(BLENDER2OPENCV, dtype=torch.float64, device=c2w.device)
    full = c2w.clone()
    full[:, :3, 3] += env_origins.double()
    vt = b2o @ torch.linalg.inv(full)       # inv(c2w_full @ b2o) = b2o @ inv(c2w_full)
    return vt.transpose(-2, -1).contiguous().float()


def c2w_from_view_matrix(view: torch.Tensor, env_origins: torch.Tensor) -> torch.Tensor:
    """The reference's own lines (env_train_gennbv.py:512-514) on whatever device `view` is on."""
    b2o = torch.tensor(BLENDER2OPENCV, dtype=torch.float32, device=view.device)
    c2w = torch.linalg.inv(view.transpose(-2, -1)) @ b2o.unsqueeze(0)
    c2w[:, :3, 3] -= env_origins
    return c2w


def render(params: torch.Tensor, poses: torch.Tensor, H: int, W: int, fov_deg: float = 90.0,
           with_rgb: bool = False):
    """Raw sensor images for N cameras (env e sees scene e % S).

    Returns depth_raw [N,H,W] float32 (negative z-depth, -inf = no hit), seg [N,H,W] int32,
    rgb [N,H,W,4] uint8 or None, c2w [N,4,4] float64 (env-local)."""
    dev = poses.device
    N, S = poses.shape[0], params.shape[0]
    prm = params.to(dev)[torch.arange(N, device=dev) % S]                # [N,4] f64
    lx, ly, hw, hr = prm[:, 0], prm[:, 1], prm[:, 2], prm[:, 3]
    K = camera_intrinsics(H, W, fov_deg).double().to(dev)
    c2w = pose_to_c2w(poses)
    u = torch.arange(W, dtype=torch.float64, device=dev)
    v = torch.arange(H, dtype=torch.float64, device=dev)
    dx = (u - K[0, 2]) / K[0, 0]
    dy = (v - K[1, 2]) / K[1, 1]
    dcam = torch.stack([dx[None, :].expand(H, W), dy[:, None].expand(H, W),
                        torch.ones(H, W, dtype=torch.float64, device=dev)], -1).view(-1, 3)  # [P,3]
    d = torch.einsum("nij,pj->npi", c2w[:, :3, :3], dcam)                # [N,P,3] world dirs, z_cam = 1
    o = c2w[:, :3, 3]                                                    # [N,3]
    a = hr / (ly / 2)
    zero, one = torch.zeros_like(lx), torch.ones_like(lx)
    # half-spaces n.x <= c : 4 walls, floor, 2 roof slopes
    nrm = torch.stack([
        torch.stack([one, zero, zero], -1), torch.stack([-one, zero, zero], -1),
        torch.stack([zero, one, zero], -1), torch.stack([zero, -one, zero], -1),
        torch.stack([zero, zero, -one], -1),
        torch.stack([zero, a, one], -1), torch.stack([zero, -a, one], -1)], 1)   # [N,7,3]
    c = torch.stack([lx / 2, lx / 2, ly / 2, ly / 2, zero, hw + hr, hw + hr], -1)   # [N,7]
    nd = torch.einsum("nki,npi->npk", nrm, d)                             # [N,P,7]
    no = torch.einsum("nki,ni->nk", nrm, o)                               # [N,7]
    t = (c - no)[:, None, :] / nd
    inf = torch.tensor(float("inf"), dtype=torch.float64, device=dev)
    t_enter = torch.where(nd < 0, t, -inf).amax(-1)
    t_exit = torch.where(nd > 0, t, inf).amin(-1)
    parallel_out = ((nd == 0) & ((no - c)[:, None, :] > 0)).any(-1)
    hit = (t_enter < t_exit) & (t_enter > 1e-3) & ~parallel_out
    # which face was entered (for shading)
    face = torch.where(nd < 0, t, -inf).argmax(-1)
    # ground plane z = 0 (segmentation id 0, env_train_base.py:26)
    tg = -o[:, None, 2] / d[..., 2]
    ground = (d[..., 2] < 0) & (tg > 1e-3)
    depth = torch.where(hit, t_enter, torch.where(ground, tg, inf))
    depth_raw = (-depth).float().view(N, H, W)                           # -inf where nothing is hit
    seg = torch.where(hit, OBJECT_SEGMENTATION_ID, 0).to(torch.int32).view(N, H, W)
    rgb = None
    if with_rgb:
        shade = torch.tensor([200, 170, 140, 110, 0, 230, 90], dtype=torch.float64, device=dev)[face]
        col = torch.where(hit, shade, torch.where(ground, 60.0, 255.0))
        rgb = torch.stack([col, col * 0.8, col * 0.6, torch.full_like(col, 255.0)], -1)
        rgb = rgb.round().to(torch.uint8).view(N, H, W, 4)
    return depth_raw, seg, rgb, c2w

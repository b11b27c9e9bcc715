"""ctypes/numpy front-end of the C oracle (oracle/gennbv_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of gennbv_oracle.c.  Importable only from
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libgennbv_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """gcc-compile the oracle in place (oracle/_build/, git-ignored)."""
    src = os.path.join(_HERE, "gennbv_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.gnbv_oracle_bresenham3d.restype = ctypes.c_int64
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def post_process_depth(raw):
    raw = _c(raw, np.float32)
    out = np.empty_like(raw)
    lib().gnbv_oracle_post_process_depth(_p(raw), _p(out), ctypes.c_int64(raw.size))
    return out


def back_projection(depth, seg, kinv, c2w):
    """-> world [N,P,3] f32, fg [N,P] bool (dense form of back_projection_fg)."""
    depth, seg = _c(depth, np.float32), _c(seg, np.int32)
    N, H, W = depth.shape
    kinv, c2w = _c(kinv, np.float32), _c(c2w, np.float32)
    world = np.empty((N, H * W, 3), np.float32)
    fg = np.empty((N, H * W), np.uint8)
    lib().gnbv_oracle_back_projection(_p(depth), _p(seg), _p(kinv), _p(c2w), N, H, W, _p(world), _p(fg))
    return world, fg.astype(bool)


def pose_to_idx(pose_xyz, range_gt, voxel_size):
    pose_xyz, range_gt, voxel_size = _c(pose_xyz, np.float32), _c(range_gt, np.float32), _c(voxel_size, np.float32)
    out = np.empty(pose_xyz.shape, np.int64)
    lib().gnbv_oracle_pose_to_idx(_p(pose_xyz), _p(range_gt), _p(voxel_size), pose_xyz.shape[0], _p(out))
    return out


def bresenham3d(src, tgt, G):
    """bresenham3D_pycuda restated: [sum_len, 3] int64, ray order, duplicates kept."""
    src, tgt = _c(src, np.int64).reshape(-1), _c(tgt, np.int64).reshape(-1, 3)
    n = lib().gnbv_oracle_bresenham3d(_p(src), _p(tgt), ctypes.c_int64(tgt.shape[0]), G, None, ctypes.c_int64(0))
    out = np.empty((n, 3), np.int64)
    lib().gnbv_oracle_bresenham3d(_p(src), _p(tgt), ctypes.c_int64(tgt.shape[0]), G, _p(out), ctypes.c_int64(n))
    return out


def voxelize_step(depth, seg, kinv, c2w, range_gt, voxel_size, pose_xyz, grid_gt, prob_grid, scanned_gt,
                  raw_depth=False, want_masks=False):
    """One env.step() of state encoding on dense fp32 grids.  prob_grid / scanned_gt [N,G,G,G] f32
    are updated IN PLACE (must be C-contiguous float32).  Returns dict(tri, cov_sum, num_targets[, masks])."""
    depth, seg = _c(depth, np.float32), _c(seg, np.int32)
    N, H, W = depth.shape
    G = grid_gt.shape[1]
    for a in (prob_grid, scanned_gt):
        assert a.dtype == np.float32 and a.flags.c_contiguous and a.shape == (N, G, G, G)
    kinv, c2w = _c(kinv, np.float32), _c(c2w, np.float32)
    range_gt, voxel_size = _c(range_gt, np.float32), _c(voxel_size, np.float32)
    pose_xyz, grid_gt = _c(pose_xyz, np.float32), _c(grid_gt, np.float32)
    tri = np.empty((N, G, G, G), np.float32)
    cov = np.empty(N, np.float32)
    nt = np.empty(N, np.int32)
    tm = np.empty((N, G, G, G), np.uint8) if want_masks else None
    rm = np.empty((N, G, G, G), np.uint8) if want_masks else None
    lib().gnbv_oracle_voxelize_step(_p(depth), _p(seg), _p(kinv), _p(c2w), _p(range_gt), _p(voxel_size),
                                    _p(pose_xyz), _p(grid_gt), _p(prob_grid), _p(scanned_gt), _p(tri),
                                    _p(cov), _p(nt), _p(tm), _p(rm), N, H, W, G, int(bool(raw_depth)))
    out = dict(tri=tri, cov_sum=cov, num_targets=nt)
    if want_masks:
        out.update(target_mask=tm.astype(bool), touched_mask=rm.astype(bool))
    return out


def gae(rewards, values, episode_starts, last_values, dones, gamma=0.99, lam=0.95):
    """[T,N] f32 arrays -> advantages, returns [T,N] f32 (buffers.py:706-724)."""
    rewards, values = _c(rewards, np.float32), _c(values, np.float32)
    T, N = rewards.shape
    es, lv, dn = _c(episode_starts, np.uint8), _c(last_values, np.float32), _c(dones, np.uint8)
    adv, ret = np.empty((T, N), np.float32), np.empty((T, N), np.float32)
    lib().gnbv_oracle_gae(_p(rewards), _p(values), _p(es), _p(lv), _p(dn), ctypes.c_double(gamma),
                          ctypes.c_double(lam), T, N, _p(adv), _p(ret))
    return adv, ret


# ---- eval accuracy (gennbv/env/env_eval_gennbv.py:250-264) ------------------------------------------------------------
def round_1cm_unique(points):
    """`torch.unique(torch.round(pts, decimals=2), dim=0)` (:254-257).  torch.round(decimals=2) is
    nearbyint(x * 100.f) / 100.f in fp32 (ATen round_decimals_kernel); unique(dim=0) returns the distinct rows in
    lexicographic order."""
    p = _c(points, np.float32).reshape(-1, 3)
    r = (np.rint(p * np.float32(100.0)) / np.float32(100.0)).astype(np.float32)
    return np.unique(r, axis=0) if r.shape[0] else r


def chamfer(x, y):
    """pytorch3d.loss.chamfer_distance with the defaults used at env_eval_gennbv.py:258 (third party, 0.7.x, not vendored in
    the reference -- its published definition: squared L2, mean over points, both directions summed), for one cloud pair,
    via an exact float64 k-d tree.  Returns (mean_x min_y d^2, mean_y min_x d^2)."""
    from scipy.spatial import cKDTree
    x, y = np.asarray(x, np.float64).reshape(-1, 3), np.asarray(y, np.float64).reshape(-1, 3)
    dx, _ = cKDTree(y).query(x, k=1)
    dy, _ = cKDTree(x).query(y, k=1)
    return float(np.mean(dx ** 2)), float(np.mean(dy ** 2))


def nn_sqdist(q, r):
    """Exact min_j |q_i - r_j|^2 in float64 (k-d tree), [P] -- checker for the per-point minima of the CUDA search."""
    from scipy.spatial import cKDTree
    d, _ = cKDTree(np.asarray(r, np.float64)).query(np.asarray(q, np.float64), k=1)
    return d ** 2

"""GPU: the reference-named free functions (gennbv_b200/utils.py) against the C oracle / recorded reference values."""
import numpy as np
import pytest
import torch

import oracle as c_oracle
from gennbv_b200 import utils
from helpers import EnvGolden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_bresenham3D_pycuda_layout_and_values():
    rng = np.random.default_rng(0)
    G = 20
    for _ in range(5):
        src = rng.integers(-30, 50, (1, 3))
        tgt = rng.integers(0, G, (int(rng.integers(1, 300)), 3))
        want = c_oracle.bresenham3d(src[0], tgt, G)
        got = utils.bresenham3D_pycuda(torch.from_numpy(src).to(DEV), torch.from_numpy(tgt).to(DEV), [G, G, G])
        assert got.dtype == torch.int64
        np.testing.assert_array_equal(got.cpu().numpy(), want)
    empty = utils.bresenham3D_pycuda(torch.tensor([[99, 99, 99]], device=DEV), torch.tensor([[99, 99, 99]], device=DEV), G)
    assert empty.shape == (0, 3)


def test_scanned_pts_and_pose_idx_match_oracle():
    g = EnvGolden("env_g20")
    t = 1
    depth = c_oracle.post_process_depth(g.depth[t])
    world, fg = c_oracle.back_projection(depth, g.seg[t], g.inv_intri, g.c2w[t])
    pts = [torch.from_numpy(world[n][fg[n]]).to(DEV) for n in range(g.N)]
    pts.append(torch.zeros(0, 3, device=DEV))                       # an env without any foreground point
    rg = torch.from_numpy(np.concatenate([g.range_gt, g.range_gt[:1]])).to(DEV)
    vs = torch.from_numpy(np.concatenate([g.voxel_size_gt, g.voxel_size_gt[:1]])).to(DEV)
    rows = utils.scanned_pts_to_idx_3D(pts, rg, vs, map_size=g.G)
    assert rows[-1] == []
    prob = np.zeros((g.N, g.G, g.G, g.G), np.float32); scan = np.zeros_like(prob)
    o = c_oracle.voxelize_step(g.depth[t], g.seg[t], g.inv_intri, g.c2w[t], g.range_gt, g.voxel_size_gt,
                               np.ascontiguousarray(g.poses[t][:, :3]), g.grid_gt, prob, scan, raw_depth=True, want_masks=True)
    for n in range(g.N):
        want = np.argwhere(o["target_mask"][n])                    # lexicographic order == torch.unique(dim=0)
        np.testing.assert_array_equal(rows[n].cpu().numpy(), want)
    poses = torch.from_numpy(np.ascontiguousarray(g.poses[t][:, :3])).to(DEV)
    idx = utils.pose_coord_to_idx_3D(poses, rg[:g.N], vs[:g.N], map_size=g.G)
    np.testing.assert_array_equal(idx.cpu().numpy(), c_oracle.pose_to_idx(g.poses[t][:, :3], g.range_gt, g.voxel_size_gt))
    tri = utils.grid_occupancy_tri_cls(torch.from_numpy(g.prob[t]).to(DEV), return_tri_cls_only=True)
    want_tri = (g.prob[t] > 0.5).astype(np.float32) - (g.prob[t] < 0).astype(np.float32)
    np.testing.assert_array_equal(tri.cpu().numpy(), want_tri)

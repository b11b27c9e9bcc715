"""PyTorch-CPU restatement of the reference's state-encoding path, op for op.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/gennbv_oracle.c header).  Where the C oracle restates the
*arithmetic*, this module restates the reference's *execution strategy* -- the same torch operators in the
same per-env Python loops -- so that bench.py's `--impl reference` arm times what "the reference's CPU
PyTorch path" costs on the host cores.  The one piece that cannot run on a CPU, the PyCUDA Bresenham
kernel (gennbv/utils.py:24-227), is served by the C restatement.

Pinned in the build container against the recorded reference roll-outs (tests/test_oracle_golden.py).
"""
import numpy as np
import torch

import oracle as c_oracle


def post_process_depth(depth_raw):
    """env_train_base.py:519-523."""
    d = torch.nan_to_num(depth_raw, neginf=0)
    d = torch.clamp(d, min=-50.0)
    return d.abs()


def back_projection_fg(depth, seg, inv_intri, c2w, pix):
    """env_train_gennbv.py:503-527 -> list of [n_i,3] world points of the foreground pixels."""
    N = depth.shape[0]
    d = depth.clone()
    fg = seg.clone() > 50
    d[~fg] = 0.0
    d = d.reshape(N, -1)
    fg = fg.reshape(N, -1)
    cp = torch.einsum("ij,jk->ijk", d, pix)
    cc = torch.einsum("ij,nkj->nki", inv_intri, cp)
    cch = torch.cat((cc, torch.ones_like(cc[..., :1])), dim=-1)
    cw = torch.einsum("nij,nkj->nki", c2w, cch)[..., :3]
    return [cw[i][fg[i]] for i in range(N)]


def scanned_pts_to_idx(pts, range_gt, vs, G):
    """gennbv/utils.py:230-270."""
    hi = range_gt[:, [0, 2, 4]] + 0.5 * vs
    lo = range_gt[:, [1, 3, 5]] - 0.5 * vs
    out = []
    for e in range(len(pts)):
        p = pts[e]
        idx = torch.floor((p - lo[e]) / vs[e]).long()
        keep = torch.all((hi[e] > p) & (p > lo[e]), dim=-1)
        v = idx[keep]
        if len(v) == 0:
            out.append([])
            continue
        v = torch.unique(v, dim=0)
        out.append(torch.clamp(v, min=0, max=G - 1))
    return out


def pose_to_idx(xyz, range_gt, vs):
    """gennbv/utils.py:273-306 (if_col=False)."""
    lo = torch.stack([range_gt[:, 1], range_gt[:, 3], range_gt[:, 5]], dim=-1) - 0.5 * vs
    return ((xyz - lo) / vs).floor().long()


def pixel_grid(H, W):
    """env_train_gennbv.py:172-181: integer pixel coordinates (u, v, 1), row-major."""
    ys, xs = torch.meshgrid(torch.linspace(0, H - 1, H), torch.linspace(0, W - 1, W), indexing="ij")
    return torch.stack([xs, ys, torch.ones_like(xs)], dim=-1).view(-1, 3)


def voxelize_step(depth_raw, seg, inv_intri, c2w, range_gt, vs, xyz, grid_gt, prob, scan, pix):
    """One update_occ_grid (env_train_gennbv.py:277-326) + coverage sum (:537); prob / scan updated in place.
    Returns (tri [N,G,G,G], cov_sum [N])."""
    N, G = prob.shape[0], prob.shape[1]
    depth = post_process_depth(depth_raw)
    pts = back_projection_fg(depth, seg, inv_intri, c2w, pix)
    idx_all = scanned_pts_to_idx(pts, range_gt, vs, G)
    src = pose_to_idx(xyz.clone(), range_gt, vs)
    occ = torch.zeros(N, G, G, G)
    for e in range(N):
        v = idx_all[e]
        if (isinstance(v, list) and len(v) == 0) or v.shape[0] == 0:
            continue
        v = torch.unique(v, dim=0, sorted=False)
        occ[e, v[:, 0], v[:, 1], v[:, 2]] = 1.0
        path = torch.from_numpy(c_oracle.bresenham3d(src[e:e + 1].int().numpy().astype(np.int64),
                                                     v.int().numpy().astype(np.int64), G))
        prob[e, path[:, 0], path[:, 1], path[:, 2]] -= 0.05
        prob[e, v[:, 0], v[:, 1], v[:, 2]] = 1.0
    tri = (prob > 0.5).to(torch.float32) - (prob < 0.0).to(torch.float32)      # utils.py:318-321
    scan.copy_(torch.clip(scan + occ * grid_gt, max=1, min=0))
    return tri, scan.sum(dim=(1, 2, 3))

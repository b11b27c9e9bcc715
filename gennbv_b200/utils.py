"""The reference's free functions (gennbv/utils.py) with their names, argument names and return types, as thin
wrappers over the kernels of libgennbv_b200 -- for callers that use them outside `Env_Train_GenNBV` (the env itself
uses the fused gnbv_voxelize_step and never materialises these index lists).  List-of-tensor return types force one
host read per env, exactly as in the reference."""
import torch

from . import _lib, ops


def _mask_to_rows(mask_words, G):
    bits = (mask_words.view(torch.uint8).unsqueeze(-1) >> torch.arange(8, device=mask_words.device, dtype=torch.uint8)) & 1
    lin = torch.nonzero(bits.reshape(-1)[: G ** 3]).flatten()
    return torch.stack([lin // (G * G), (lin // G) % G, lin % G], dim=-1)


def scanned_pts_to_idx_3D(pts_target, range_gt, voxel_size_gt, map_size=256):
    """gennbv/utils.py:230-270 -> list of [m_i,3] int64 (unique, sorted, clamped) or [] per env."""
    out = []
    G, L = int(map_size), _lib.lib()
    words = (G ** 3 + 31) // 32
    for e, pts in enumerate(pts_target):
        pts = pts.contiguous().float()
        if not pts.is_cuda:
            raise RuntimeError("scanned_pts_to_idx_3D: expected CUDA tensors (no CPU path)")
        mask = torch.zeros(words, dtype=torch.int32, device=pts.device)
        rg, vs = range_gt[e].contiguous(), voxel_size_gt[e].contiguous()
        _lib.check(L.gnbv_points_to_voxel_mask(pts.data_ptr(), pts.shape[0], rg.data_ptr(), vs.data_ptr(), mask.data_ptr(), G,
                                               ops._stream()), "gnbv_points_to_voxel_mask")
        rows = _mask_to_rows(mask, G)
        out.append(rows if rows.shape[0] else [])
    return out


def pose_coord_to_idx_3D(poses, range_gt, voxel_size_gt, map_size=256, if_col=False):
    """gennbv/utils.py:273-306: unclamped camera voxel (elementwise IEEE fp32 ops: identical on any device)."""
    lo = torch.stack([range_gt[:, 1], range_gt[:, 3], range_gt[:, 5]], dim=-1) - 0.5 * voxel_size_gt
    assert poses.shape[1] == 3, f"Invalid poses shape: {poses.shape}"
    idx = ((poses - lo) / voxel_size_gt).floor().long()
    if if_col:
        idx[(idx < 0).any(dim=-1)] = -1
        idx[(idx > map_size - 1).any(dim=-1)] = -1
    return idx


def bresenham3D_pycuda(pts_source, pts_target, map_size):
    """gennbv/utils.py:24-227 -> [sum_len,3] int64: in-bounds voxels of every ray, ray order, duplicates kept.
    (One pre-compiled sm_100a kernel on torch's stream instead of a PyCUDA JIT build per call.)"""
    if isinstance(map_size, list):
        assert len(map_size) == 3 and map_size[0] == map_size[1] == map_size[2], "map_size must be cubic"
        map_size = map_size[0]
    if not pts_target.is_cuda:
        raise RuntimeError("bresenham3D_pycuda: expected CUDA tensors (no CPU path)")
    src = pts_source.int().contiguous().view(-1)
    tgt = pts_target.int().contiguous()
    n, L, s = tgt.shape[0], _lib.lib(), ops._stream()
    counts = torch.zeros(n, dtype=torch.int64, device=tgt.device)
    _lib.check(L.gnbv_bresenham_rays(src.data_ptr(), tgt.data_ptr(), n, int(map_size), counts.data_ptr(), None, None, s),
               "gnbv_bresenham_rays")
    offsets = torch.cumsum(counts, 0) - counts
    total = int(counts.sum()) if n else 0
    out = torch.empty(total, 3, dtype=torch.int64, device=tgt.device)
    if total:
        _lib.check(L.gnbv_bresenham_rays(src.data_ptr(), tgt.data_ptr(), n, int(map_size), None, offsets.data_ptr(),
                                         out.data_ptr(), s), "gnbv_bresenham_rays")
    return out


def grid_occupancy_tri_cls(grid_prob, threshold_occu=0.5, threshold_free=0.0, return_tri_cls_only=False):
    """gennbv/utils.py:309-325 (three comparisons; inside the env this is fused into the grid-update kernel)."""
    occ = (grid_prob > threshold_occu).to(torch.float32)
    tri = occ - (grid_prob < threshold_free).to(torch.float32)
    return tri if return_tri_cls_only else (occ, tri)

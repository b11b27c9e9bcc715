"""Thin torch-tensor front-end of the C ABI: raw device pointers + the current CUDA stream.

PyTorch is plumbing here (device memory, streams); every computation is a kernel of
libgennbv_b200.so.  All functions raise if a tensor is not a contiguous CUDA tensor of the
expected dtype -- nothing is silently copied, cast or run on the CPU.
"""
import ctypes

import torch

from . import _lib

GNBV_RAW_DEPTH = 1


def _ptr(t, dtype, name, shape=None):
    if t is None:
        return None
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (gennbv_b200 has no CPU path)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: expected a contiguous tensor")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise RuntimeError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def voxelize_workspace(num_envs, grid_size, device):
    nbytes = _lib.lib().gnbv_voxelize_workspace_bytes(num_envs, grid_size)
    return torch.empty(nbytes, dtype=torch.uint8, device=device)


def voxelize_step(depth, seg, kinv, c2w, range_gt, voxel_size, pose_xyz, grid_gt, prob_grid, scanned_gt,
                  tri_out, cov_sum, num_targets=None, workspace=None, raw_depth=False, tri_row_stride=None):
    """gnbv_voxelize_step (include/gennbv_b200.h).  prob_grid / scanned_gt are updated in place;
    tri_out may be a dense [N,G,G,G] tensor or the grid columns of the flat observation buffer
    (pass tri_row_stride = row length in elements and tri_out = the 1-D view starting at the first grid element)."""
    N, H, W = depth.shape
    G = grid_gt.shape[1]
    V = G * G * G
    if workspace is None:
        workspace = voxelize_workspace(N, G, depth.device)
    if tri_row_stride is None:
        tri_row_stride = V
        tri_ptr = _ptr(tri_out, torch.float32, "tri_out", (N, G, G, G))
    else:
        tri_ptr = _ptr(tri_out, torch.float32, "tri_out")
        if tri_out.numel() < (N - 1) * tri_row_stride + V:
            raise RuntimeError("tri_out view too small for tri_row_stride")
    rc = _lib.lib().gnbv_voxelize_step(
        _ptr(depth, torch.float32, "depth", (N, H, W)), _ptr(seg, torch.int32, "seg", (N, H, W)),
        _ptr(kinv, torch.float32, "kinv", (3, 3)), _ptr(c2w, torch.float32, "c2w", (N, 4, 4)),
        _ptr(range_gt, torch.float32, "range_gt", (N, 6)), _ptr(voxel_size, torch.float32, "voxel_size", (N, 3)),
        _ptr(pose_xyz, torch.float32, "pose_xyz", (N, 3)), _ptr(grid_gt, torch.float32, "grid_gt", (N, G, G, G)),
        _ptr(prob_grid, torch.float32, "prob_grid", (N, G, G, G)),
        _ptr(scanned_gt, torch.float32, "scanned_gt", (N, G, G, G)),
        tri_ptr, tri_row_stride, _ptr(cov_sum, torch.float32, "cov_sum", (N,)),
        _ptr(num_targets, torch.int32, "num_targets", (N,)),
        _ptr(workspace, torch.uint8, "workspace"), workspace.numel(), N, H, W, G,
        GNBV_RAW_DEPTH if raw_depth else 0, _stream())
    _lib.check(rc, "gnbv_voxelize_step")
    return workspace


def scan_raycast(depth, seg, kinv, c2w, range_gt, voxel_size, pose_xyz, grid_size, workspace, num_targets=None,
                 raw_depth=False):
    """gnbv_scan_raycast: phase 1 of the step (masks land in `workspace`)."""
    N, H, W = depth.shape
    rc = _lib.lib().gnbv_scan_raycast(
        _ptr(depth, torch.float32, "depth", (N, H, W)), _ptr(seg, torch.int32, "seg", (N, H, W)),
        _ptr(kinv, torch.float32, "kinv", (3, 3)), _ptr(c2w, torch.float32, "c2w", (N, 4, 4)),
        _ptr(range_gt, torch.float32, "range_gt", (N, 6)), _ptr(voxel_size, torch.float32, "voxel_size", (N, 3)),
        _ptr(pose_xyz, torch.float32, "pose_xyz", (N, 3)), _ptr(num_targets, torch.int32, "num_targets", (N,)),
        _ptr(workspace, torch.uint8, "workspace"), workspace.numel(), N, H, W, grid_size,
        GNBV_RAW_DEPTH if raw_depth else 0, _stream())
    _lib.check(rc, "gnbv_scan_raycast")


def grid_update(grid_gt, prob_grid, scanned_gt, tri_out, cov_sum, workspace, tri_row_stride=None, sparse=False):
    """gnbv_grid_update: phase 2 of the step (dense prob / tri / scanned_gt / coverage pass).  sparse=True:
    gnbv_grid_update_sparse (untouched 16-byte groups are not rewritten; cov_sum is in/out and incremented)."""
    N, G = grid_gt.shape[0], grid_gt.shape[1]
    V = G ** 3
    if tri_row_stride is None:
        tri_row_stride, tri_ptr = V, _ptr(tri_out, torch.float32, "tri_out", (N, G, G, G))
    else:
        tri_ptr = _ptr(tri_out, torch.float32, "tri_out")
        if tri_out.numel() < (N - 1) * tri_row_stride + V:
            raise RuntimeError("tri_out view too small for tri_row_stride")
    fn = _lib.lib().gnbv_grid_update_sparse if sparse else _lib.lib().gnbv_grid_update
    rc = fn(
        _ptr(grid_gt, torch.float32, "grid_gt", (N, G, G, G)), _ptr(prob_grid, torch.float32, "prob_grid", (N, G, G, G)),
        _ptr(scanned_gt, torch.float32, "scanned_gt", (N, G, G, G)), tri_ptr, tri_row_stride,
        _ptr(cov_sum, torch.float32, "cov_sum", (N,)), _ptr(workspace, torch.uint8, "workspace"), workspace.numel(),
        N, G, _stream())
    _lib.check(rc, "gnbv_grid_update")


def voxelize_masks(workspace, num_envs, grid_size):
    """Boolean [N,G,G,G] target / touched masks of the last voxelize_step on `workspace` (debug / compat path)."""
    t, r, w = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int64()
    rc = _lib.lib().gnbv_voxelize_masks(workspace.data_ptr(), num_envs, grid_size, ctypes.byref(t), ctypes.byref(r),
                                        ctypes.byref(w))
    _lib.check(rc, "gnbv_voxelize_masks")
    words, V = w.value, grid_size ** 3
    base = workspace.data_ptr()
    out = []
    for p in (t.value, r.value):
        off = p - base
        u8 = workspace[off: off + num_envs * words * 4].view(num_envs, words * 4)
        bits = (u8.unsqueeze(-1) >> torch.arange(8, device=u8.device, dtype=torch.uint8)) & 1
        out.append(bits.view(num_envs, -1)[:, :V].bool().view(num_envs, grid_size, grid_size, grid_size))
    return out[0], out[1]


def reset_grids(prob_grid, scanned_gt, reset_flags):
    N, G = prob_grid.shape[0], prob_grid.shape[1]
    rc = _lib.lib().gnbv_reset_grids(_ptr(prob_grid, torch.float32, "prob_grid", (N, G, G, G)),
                                     _ptr(scanned_gt, torch.float32, "scanned_gt", (N, G, G, G)),
                                     _ptr(reset_flags, torch.uint8, "reset_flags", (N,)), N, G, _stream())
    _lib.check(rc, "gnbv_reset_grids")


def gae(rewards, values, episode_starts, last_values, dones, gamma, gae_lambda, advantages, returns):
    T, N = rewards.shape
    rc = _lib.lib().gnbv_gae(_ptr(rewards, torch.float32, "rewards", (T, N)), _ptr(values, torch.float32, "values", (T, N)),
                             _ptr(episode_starts, torch.uint8, "episode_starts", (T, N)),
                             _ptr(last_values, torch.float32, "last_values", (N,)), _ptr(dones, torch.uint8, "dones", (N,)),
                             float(gamma), float(gae_lambda), T, N, _ptr(advantages, torch.float32, "advantages", (T, N)),
                             _ptr(returns, torch.float32, "returns", (T, N)), _stream())
    _lib.check(rc, "gnbv_gae")


def sgemm(A, a_strides, B, b_strides, C, M, N, K, bias=None, relu=False, ldc=None):
    """gnbv_sgemm: C[M,N] = relu?(A*B + bias); a_strides = (stride_m, stride_k) of A, b_strides = (stride_k, stride_n) of B
    in elements."""
    L = _lib.lib()
    nbytes = L.gnbv_sgemm_workspace_bytes(M, N, K)
    ws = torch.empty(max(nbytes, 4), dtype=torch.uint8, device=C.device)
    rc = L.gnbv_sgemm(_ptr(A, torch.float32, "A"), a_strides[0], a_strides[1], _ptr(B, torch.float32, "B"), b_strides[0],
                      b_strides[1], _ptr(C, torch.float32, "C"), N if ldc is None else ldc, M, N, K,
                      _ptr(bias, torch.float32, "bias"), int(relu), ws.data_ptr() if nbytes else None, nbytes, _stream())
    _lib.check(rc, "gnbv_sgemm")
    return C


def tc_gemm(A, a_strides, B, b_strides, C, M, N, K, bias=None, relu=False, ldc=None, workspace=None):
    """gnbv_tc_gemm: the tcgen05 / 3xTF32 version of sgemm.  Returns the workspace (float slot 0 = sticky error flag)."""
    L = _lib.lib()
    nbytes = L.gnbv_tc_gemm_workspace_bytes(M, N, K)
    if workspace is None or workspace.numel() * 4 < nbytes:
        workspace = torch.zeros(nbytes // 4, dtype=torch.float32, device=C.device)
    rc = L.gnbv_tc_gemm(_ptr(A, torch.float32, "A"), a_strides[0], a_strides[1], _ptr(B, torch.float32, "B"), b_strides[0],
                        b_strides[1], _ptr(C, torch.float32, "C"), N if ldc is None else ldc, M, N, K,
                        _ptr(bias, torch.float32, "bias"), int(relu), workspace.data_ptr(), workspace.numel() * 4, _stream())
    _lib.check(rc, "gnbv_tc_gemm")
    return workspace


def sort_unique(keys, key_bits=64):
    """gnbv_sort_unique_u64: the distinct values of an int64 CUDA tensor in ascending order (in-tree radix sort + compaction;
    what `torch.unique` did on the eval path).  `keys` must be non-negative below 2**key_bits; it is used as scratch."""
    n = keys.numel()
    if n == 0:
        return keys.new_empty(0)
    L = _lib.lib()
    nbytes = L.gnbv_sort_unique_workspace_bytes(n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=keys.device)
    out = torch.empty(n, dtype=torch.int64, device=keys.device)
    count = torch.zeros(1, dtype=torch.int64, device=keys.device)
    rc = L.gnbv_sort_unique_u64(_ptr(keys, torch.int64, "keys"), n, int(key_bits), out.data_ptr(), count.data_ptr(), ws.data_ptr(),
                                nbytes, _stream())
    _lib.check(rc, "gnbv_sort_unique_u64")
    return out[:int(count)]


def points_to_keys(points):
    """gnbv_points_to_keys: [n,3] f32 -> packed 1 cm lattice keys (ascending key order == lexicographic row order)."""
    n = points.shape[0]
    keys = torch.empty(n, dtype=torch.int64, device=points.device)
    if n:
        _lib.check(_lib.lib().gnbv_points_to_keys(_ptr(points, torch.float32, "points", (n, 3)), n, keys.data_ptr(), _stream()),
                   "gnbv_points_to_keys")
    return keys


def keys_to_points(keys):
    pts = torch.empty(keys.shape[0], 3, device=keys.device)
    if keys.shape[0]:
        _lib.check(_lib.lib().gnbv_keys_to_points(_ptr(keys, torch.int64, "keys"), keys.shape[0], pts.data_ptr(), _stream()),
                   "gnbv_keys_to_points")
    return pts

"""Eval accuracy of the reference's eval env (gennbv/env/env_eval_gennbv.py:160-164, 253-264): scanned-point history,
1 cm dedup and the chamfer distance to the GT cloud, with `pytorch3d.loss.chamfer_distance` replaced by the exact 1-NN
kernels of libgennbv_b200 (gnbv_chamfer).  pytorch3d is not vendored by the reference and absent here: its published
definition is restated (squared L2, mean over points, both directions summed) -- parity unpinned."""
import torch

from . import _lib, ops


GRID_MIN_PAIRS = 1 << 24          # below this many point pairs per cloud the P1*P2 scan is as fast as building a grid
GRID_WORKSPACE_LIMIT = 2 << 30    # bytes of grid workspace per launch group; larger batches are processed in chunks of clouds


def chamfer_distance(x, y, method="auto"):
    """x [1,P1,3] / [P1,3] or a list of [P_e,3] clouds; y likewise -> (loss, None) like pytorch3d (loss averaged over the
    batch, `batch_reduction="mean"`).  All tensors must be float32 CUDA tensors."""
    xs = [x.reshape(-1, 3)] if isinstance(x, torch.Tensor) and x.dim() <= 2 else ([c.reshape(-1, 3) for c in x])
    ys = [y.reshape(-1, 3)] if isinstance(y, torch.Tensor) and y.dim() <= 2 else ([c.reshape(-1, 3) for c in y])
    if len(xs) != len(ys):
        raise ValueError("x and y must hold the same number of clouds")
    cx, cy = chamfer_terms(xs, ys, method=method)
    return (cx + cy).mean(), None


def _pack(cs, dev):
    sizes = [int(c.shape[0]) for c in cs]
    off = [0]
    for n in sizes:
        off.append(off[-1] + n)
    pts = torch.cat([c.float().contiguous() for c in cs], 0).contiguous() if off[-1] > 0 else torch.zeros(0, 3, device=dev)
    return pts, torch.tensor(off, dtype=torch.int64, device=dev), sizes


def default_cells_per_axis(n_ref):
    """Cells along the longest box axis for a reference cloud of n_ref surface points: about two points per occupied
    cell for a closed surface (~6 C^2 occupied cells), capped so that the cell array stays at 8 MB per cloud."""
    return max(4, min(128, int(round((max(n_ref, 1) / 12.0) ** 0.5))))


def chamfer_terms(xs, ys, method="auto", cells_per_axis=None, return_min=False):
    """Per-cloud one-directional terms (mean_i min_j d^2, mean_j min_i d^2) as two [E] tensors, for lists of clouds.

    method: "brute" (P1*P2 scan, gnbv_chamfer), "grid" (exact uniform-grid search, gnbv_chamfer_grid) or "auto"
    (grid once a cloud pair has >= GRID_MIN_PAIRS pairs).  With return_min=True the per-point minima (two packed
    [sum P] tensors) are returned as well."""
    dev = xs[0].device
    if dev.type != "cuda":
        raise RuntimeError("chamfer_distance: expected CUDA tensors (no CPU path)")
    xp, _, nx = _pack(xs, dev)
    yp, _, ny = _pack(ys, dev)
    return chamfer_terms_packed(xp, nx, yp, ny, method=method, cells_per_axis=cells_per_axis, return_min=return_min)


def chamfer_terms_packed(xp, nx, yp, ny, method="auto", cells_per_axis=None, return_min=False):
    """Same for clouds that are already packed: xp [sum nx, 3] / yp [sum ny, 3] float32 CUDA tensors, nx / ny the per-cloud
    point counts (host lists)."""
    E, dev = len(nx), xp.device
    if dev.type != "cuda":
        raise RuntimeError("chamfer_distance: expected CUDA tensors (no CPU path)")
    if len(ny) != E or xp.shape[0] != sum(nx) or yp.shape[0] != sum(ny):
        raise ValueError("packed clouds and their sizes do not match")
    if method not in ("auto", "brute", "grid"):
        raise ValueError(f"unknown method {method!r}")
    L, s = _lib.lib(), ops._stream()
    xp, yp = xp.float().contiguous(), yp.float().contiguous()
    offs = lambda n: torch.tensor([0] + list(torch.tensor(n, dtype=torch.int64).cumsum(0)), dtype=torch.int64, device=dev)
    xo, yo = offs(nx), offs(ny)
    if method == "auto":
        method = "grid" if max(a * b for a, b in zip(nx, ny)) >= GRID_MIN_PAIRS else "brute"
    cx, cy = torch.empty(E, device=dev), torch.empty(E, device=dev)
    min_x = torch.empty(xp.shape[0], device=dev) if return_min else None
    min_y = torch.empty(yp.shape[0], device=dev) if return_min else None
    if method == "brute":
        ws = torch.empty(L.gnbv_chamfer_workspace_bytes(E) // 4, device=dev)
        _lib.check(L.gnbv_chamfer(xp.data_ptr(), xo.data_ptr(), yp.data_ptr(), yo.data_ptr(), E, cx.data_ptr(), cy.data_ptr(),
                                  ws.data_ptr(), ws.numel() * 4, s), "gnbv_chamfer")
        if return_min:
            for q, qo, r, ro, out in ((xp, xo, yp, yo, min_x), (yp, yo, xp, xo, min_y)):
                _lib.check(L.gnbv_nn_sqdist_brute(q.data_ptr(), qo.data_ptr(), r.data_ptr(), ro.data_ptr(), E, out.data_ptr(),
                                                  ws.data_ptr(), ws.numel() * 4, s), "gnbv_nn_sqdist_brute")
    else:
        C = int(cells_per_axis) if cells_per_axis is not None else default_cells_per_axis(max(max(nx), max(ny)))
        e0 = 0
        while e0 < E:                                   # chunks of clouds bounded by the workspace limit
            e1 = e0 + 1
            while e1 < E and L.gnbv_chamfer_grid_workspace_bytes(e1 + 1 - e0, sum(nx[e0:e1 + 1]), sum(ny[e0:e1 + 1]), C) \
                    <= GRID_WORKSPACE_LIMIT:
                e1 += 1
            tx, ty = sum(nx[e0:e1]), sum(ny[e0:e1])
            x0, y0 = sum(nx[:e0]), sum(ny[:e0])
            xo_c, yo_c = (xo[e0:e1 + 1] - x0).contiguous(), (yo[e0:e1 + 1] - y0).contiguous()
            nbytes = L.gnbv_chamfer_grid_workspace_bytes(e1 - e0, tx, ty, C)
            if nbytes == 0:
                raise RuntimeError("gnbv_chamfer_grid_workspace_bytes rejected the arguments")
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)     # torch allocations are 512-byte aligned
            f32 = 4
            _lib.check(L.gnbv_chamfer_grid(
                xp.data_ptr() + x0 * 3 * f32, xo_c.data_ptr(), yp.data_ptr() + y0 * 3 * f32, yo_c.data_ptr(), e1 - e0, tx, ty, C,
                cx.data_ptr() + e0 * f32, cy.data_ptr() + e0 * f32,
                None if min_x is None else min_x.data_ptr() + x0 * f32, None if min_y is None else min_y.data_ptr() + y0 * f32,
                ws.data_ptr(), nbytes, s), "gnbv_chamfer_grid")
            e0 = e1
    return (cx, cy, min_x, min_y) if return_min else (cx, cy)


def accuracy_from_history(pts_history, pc_gt):
    """env_eval_gennbv.py:253-261 for one env: 1 cm voxel dedup of the scanned points, then chamfer to the GT cloud.
    Note the reference writes `(chamfer_distance(...) * 100)[0]`: the tuple is repeated 100 times and `[0]` is the
    unscaled loss -- reproduced."""
    from . import ops
    pc = ops.keys_to_points(ops.sort_unique(ops.points_to_keys(pts_history.contiguous()), key_bits=54))      # 1 cm lattice dedup
    return chamfer_distance(pc.unsqueeze(0), pc_gt.unsqueeze(0))[0].unsqueeze(0)

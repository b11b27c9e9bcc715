"""Eval accuracy of the reference's eval env (gennbv/env/env_eval_gennbv.py:160-164, 253-264): scanned-point history,
1 cm dedup and the chamfer distance to the GT cloud, with `pytorch3d.loss.chamfer_distance` replaced by the exact 1-NN
kernels of libgennbv_b200 (gnbv_chamfer).  pytorch3d is not vendored by the reference and absent here: its published
definition is restated (squared L2, mean over points, both directions summed) -- parity unpinned."""
import torch

from . import _lib, ops


def chamfer_distance(x, y):
    """x [1,P1,3] / [P1,3] or a list of [P_e,3] clouds; y likewise -> (loss, None) like pytorch3d (loss averaged over the
    batch, `batch_reduction="mean"`).  All tensors must be float32 CUDA tensors."""
    xs = [x.reshape(-1, 3)] if isinstance(x, torch.Tensor) and x.dim() <= 2 else ([c.reshape(-1, 3) for c in x])
    ys = [y.reshape(-1, 3)] if isinstance(y, torch.Tensor) and y.dim() <= 2 else ([c.reshape(-1, 3) for c in y])
    if len(xs) != len(ys):
        raise ValueError("x and y must hold the same number of clouds")
    cx, cy = chamfer_terms(xs, ys)
    return (cx + cy).mean(), None


def chamfer_terms(xs, ys):
    """Per-cloud one-directional terms (mean_i min_j d^2, mean_j min_i d^2) as two [E] tensors."""
    E, dev = len(xs), xs[0].device
    if dev.type != "cuda":
        raise RuntimeError("chamfer_distance: expected CUDA tensors (no CPU path)")
    pack = lambda cs: (torch.cat([c.float().contiguous() for c in cs], 0).contiguous(),
                       torch.tensor([0] + list(torch.tensor([c.shape[0] for c in cs]).cumsum(0)), dtype=torch.int64, device=dev))
    xp, xo = pack(xs)
    yp, yo = pack(ys)
    L = _lib.lib()
    ws = torch.empty(L.gnbv_chamfer_workspace_bytes(E) // 4, device=dev)
    cx, cy = torch.empty(E, device=dev), torch.empty(E, device=dev)
    _lib.check(L.gnbv_chamfer(xp.data_ptr(), xo.data_ptr(), yp.data_ptr(), yo.data_ptr(), E, cx.data_ptr(), cy.data_ptr(),
                              ws.data_ptr(), ws.numel() * 4, ops._stream()), "gnbv_chamfer")
    return cx, cy


def accuracy_from_history(pts_history, pc_gt):
    """env_eval_gennbv.py:253-261 for one env: 1 cm voxel dedup of the scanned points, then chamfer to the GT cloud.
    Note the reference writes `(chamfer_distance(...) * 100)[0]`: the tuple is repeated 100 times and `[0]` is the
    unscaled loss -- reproduced."""
    pc = torch.unique(torch.round(pts_history, decimals=2), dim=0)
    return chamfer_distance(pc.unsqueeze(0), pc_gt.unsqueeze(0))[0].unsqueeze(0)

"""GPU: exact 1-NN chamfer kernels against a float64 brute force (torch.cdist), and the uniform-grid search
against the scan kernel (bit-identical per-point minima) and a float64 k-d tree."""
import numpy as np
import pytest
import torch

from gennbv_b200 import chamfer

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def brute(x, y):
    d = torch.cdist(x.double(), y.double()) ** 2
    return d.min(1).values.mean() + d.min(0).values.mean()


@pytest.mark.parametrize("P1,P2,seed", [(1, 1, 0), (257, 2049, 1), (5000, 10000, 2), (3, 4100, 3)])
def test_chamfer_matches_bruteforce(P1, P2, seed):
    g = torch.Generator().manual_seed(seed)
    x, y = torch.randn(P1, 3, generator=g) * 2, torch.randn(P2, 3, generator=g) * 2 + 0.1
    loss, normals = chamfer.chamfer_distance(x.to(DEV).unsqueeze(0), y.to(DEV).unsqueeze(0))
    assert normals is None
    assert abs(float(loss) - float(brute(x, y))) <= 1e-5 * float(brute(x, y)) + 1e-9


def test_batched_ragged_clouds_and_properties():
    g = torch.Generator().manual_seed(7)
    xs = [torch.randn(n, 3, generator=g) for n in (100, 1, 3000)]
    ys = [torch.randn(n, 3, generator=g) for n in (50, 700, 2)]
    cx, cy = chamfer.chamfer_terms([c.to(DEV) for c in xs], [c.to(DEV) for c in ys])
    for e in range(3):
        d = torch.cdist(xs[e].double(), ys[e].double()) ** 2
        assert abs(float(cx[e]) - float(d.min(1).values.mean())) < 1e-5 * float(d.min(1).values.mean()) + 1e-9
        assert abs(float(cy[e]) - float(d.min(0).values.mean())) < 1e-5 * float(d.min(0).values.mean()) + 1e-9
    # identical clouds -> 0; symmetry of the summed loss
    a, b = xs[2].to(DEV), ys[0].to(DEV)
    assert float(chamfer.chamfer_distance(a, a)[0]) == 0.0
    assert float(chamfer.chamfer_distance(a, b)[0]) == float(chamfer.chamfer_distance(b, a)[0])


def test_accuracy_from_history_dedups_at_1cm():
    g = torch.Generator().manual_seed(3)
    gt = torch.rand(2000, 3, generator=g)
    hist = torch.cat([gt[:500] + 0.001, gt[:500] + 0.002, gt[500:900]], 0)       # duplicates within 1 cm
    acc = chamfer.accuracy_from_history(hist.to(DEV), gt.to(DEV))
    pc = torch.unique(torch.round(hist, decimals=2), dim=0)
    assert acc.shape == (1,) and abs(float(acc) - float(brute(pc, gt))) < 1e-5 * float(brute(pc, gt))


# ---- exact uniform-grid search (gnbv_chamfer_grid) -------------------------------------------------------------------------
def _surface(n, seed, jitter=0.0, size=(5.3, 4.1, 3.7)):
    """points on the faces of a box (a house-like closed surface), optionally snapped to the eval env's 1 cm lattice"""
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(n, 3, generator=g)
    face = torch.randint(0, 6, (n,), generator=g)
    p = u.clone()
    for a in range(3):
        p[face == 2 * a, a] = 0.0
        p[face == 2 * a + 1, a] = 1.0
    p = p * torch.tensor(size) - torch.tensor([size[0] / 2, size[1] / 2, 0.0])
    return p + jitter * (torch.rand(n, 3, generator=g) - 0.5)


@pytest.mark.parametrize("cells", [None, 1, 7, 160])
def test_grid_search_minima_are_bit_identical_to_the_scan(cells):
    scan = torch.round(_surface(60000, 1, jitter=0.02), decimals=2)
    gt = _surface(40000, 2)
    xs, ys = [scan.to(DEV), gt[:5000].to(DEV), gt[:1].to(DEV)], [gt.to(DEV), scan[:3000].to(DEV), scan[:777].to(DEV)]
    bx, by, bmx, bmy = chamfer.chamfer_terms(xs, ys, method="brute", return_min=True)
    gx, gy, gmx, gmy = chamfer.chamfer_terms(xs, ys, method="grid", cells_per_axis=cells, return_min=True)
    assert torch.equal(bmx, gmx) and torch.equal(bmy, gmy)
    assert torch.equal(bx, gx) and torch.equal(by, gy)            # same block decomposition -> same summation order


def test_grid_search_far_queries_sparse_and_degenerate_clouds():
    import oracle as c_oracle
    g = torch.Generator().manual_seed(5)
    far = torch.randn(3000, 3, generator=g) * 200 + 30          # queries far outside the reference box (escape path)
    vol = torch.randn(8000, 3, generator=g)
    plane = torch.cat([torch.rand(4000, 2, generator=g) * 3, torch.full((4000, 1), 0.25)], 1)
    same = torch.tensor([[1.0, 2.0, 3.0]]).repeat(50, 1)
    xs, ys = [far, vol, vol[:2000], vol], [vol, plane, same, vol]
    d = lambda cs: [c.to(DEV) for c in cs]
    bx, by, bmx, bmy = chamfer.chamfer_terms(d(xs), d(ys), method="brute", return_min=True)
    gx, gy, gmx, gmy = chamfer.chamfer_terms(d(xs), d(ys), method="grid", cells_per_axis=40, return_min=True)
    assert torch.equal(bmx, gmx) and torch.equal(bmy, gmy)
    assert float(gx[3]) == 0.0 and float(gy[3]) == 0.0          # identical clouds
    # independent check of the minima: float64 k-d tree (oracle)
    want = c_oracle.nn_sqdist(xs[0].numpy(), ys[0].numpy())
    np.testing.assert_allclose(gmx[:3000].cpu().numpy(), want, rtol=2e-5)


def test_grid_path_chunks_large_batches(monkeypatch):
    xs = [_surface(3000 + 100 * e, 10 + e).to(DEV) for e in range(6)]
    ys = [_surface(2000 + 50 * e, 30 + e).to(DEV) for e in range(6)]
    whole = chamfer.chamfer_terms(xs, ys, method="grid", cells_per_axis=32, return_min=True)
    monkeypatch.setattr(chamfer, "GRID_WORKSPACE_LIMIT", 1 << 20)          # forces chunks of one or two clouds
    parts = chamfer.chamfer_terms(xs, ys, method="grid", cells_per_axis=32, return_min=True)
    for a, b in zip(whole, parts):
        assert torch.equal(a, b)


def test_auto_method_uses_the_grid_for_large_pairs_and_matches_float64():
    x, y = torch.round(_surface(30000, 3, jitter=0.03), decimals=2), _surface(20000, 4)
    assert x.shape[0] * y.shape[0] >= chamfer.GRID_MIN_PAIRS
    loss, _ = chamfer.chamfer_distance(x.to(DEV), y.to(DEV))
    import oracle as c_oracle
    cx, cy = c_oracle.chamfer(x.numpy(), y.numpy())
    assert abs(float(loss) - (cx + cy)) <= 1e-5 * (cx + cy)

"""GPU: exact 1-NN chamfer kernels against a float64 brute force (torch.cdist)."""
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

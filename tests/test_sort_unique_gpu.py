"""GPU: the in-tree radix sort + unique (gnbv_sort_unique_u64) against torch.unique / numpy on the host, and the key packing of
arbitrary points (gnbv_points_to_keys) against `torch.unique(torch.round(pts, decimals=2), dim=0)` (env_eval_gennbv.py:254-257)."""
import numpy as np
import pytest
import torch

from gennbv_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("n,bits,distinct", [(1, 54, 1), (31, 8, 5), (4096, 54, 300), (4097, 63, 4097), (100_003, 54, 20_000),
                                              (3_000_000, 62, 900_000)])
def test_sort_unique_matches_numpy(n, bits, distinct):
    rng = np.random.default_rng(n)
    pool = rng.integers(0, 2 ** bits, size=distinct, dtype=np.int64)
    keys = pool[rng.integers(0, distinct, size=n)]
    want = np.unique(keys)
    got = ops.sort_unique(torch.from_numpy(keys).to(DEV), key_bits=bits)
    np.testing.assert_array_equal(got.cpu().numpy(), want)


def test_edge_values_and_full_width():
    keys = torch.tensor([0, 2 ** 63 - 1, 5, 5, 0, 2 ** 62, 2 ** 63 - 1, 1], dtype=torch.int64, device=DEV)
    got = ops.sort_unique(keys.clone(), key_bits=64)
    assert got.tolist() == [0, 1, 5, 2 ** 62, 2 ** 63 - 1]
    assert ops.sort_unique(torch.empty(0, dtype=torch.int64, device=DEV)).numel() == 0


def test_point_dedup_equals_torch_round_unique():
    g = torch.Generator().manual_seed(0)
    pts = (torch.rand(50_000, 3, generator=g) - 0.5) * 3.0
    pts = torch.cat([pts, pts[:7000] + 0.001, torch.tensor([[0.005, -0.005, 0.015], [-1.2345, 2.0, 0.0]])])     # near-duplicates, ties
    want = torch.unique(torch.round(pts, decimals=2), dim=0)
    got = ops.keys_to_points(ops.sort_unique(ops.points_to_keys(pts.to(DEV)), key_bits=54)).cpu()
    assert got.shape == want.shape
    assert torch.equal(got, want + 0.0)            # (+0.0: the reference may hold -0.0 where the lattice decode gives +0.0)

"""GPU: gnbv_gae against the C oracle (bit-exact: same fp32 operation order as buffers.py:706-724)."""
import numpy as np
import pytest
import torch

import oracle as c_oracle
from gennbv_b200 import ops

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("T,N,seed", [(128, 256, 0), (1, 1, 1), (7, 130, 2), (128, 2048, 3)])
def test_gae_bit_exact(T, N, seed):
    rng = np.random.default_rng(seed)
    r = rng.standard_normal((T, N)).astype(np.float32) * 3
    v = rng.standard_normal((T, N)).astype(np.float32)
    es = (rng.random((T, N)) < 0.05).astype(np.uint8)
    lv = rng.standard_normal(N).astype(np.float32)
    dn = (rng.random(N) < 0.3).astype(np.uint8)
    adv_o, ret_o = c_oracle.gae(r, v, es, lv, dn, 0.99, 0.95)
    d = lambda a: torch.from_numpy(a).cuda()
    adv, ret = torch.empty(T, N, device="cuda"), torch.empty(T, N, device="cuda")
    ops.gae(d(r), d(v), d(es), d(lv), d(dn), 0.99, 0.95, adv, ret)
    np.testing.assert_array_equal(adv.cpu().numpy(), adv_o)
    np.testing.assert_array_equal(ret.cpu().numpy(), ret_o)

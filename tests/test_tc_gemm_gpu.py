"""GPU: the tcgen05 / 3xTF32 GEMM (gnbv_tc_gemm) against a float64 product: fp32-grade accuracy on tensor cores."""
import pytest
import torch

from gennbv_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("M,N,K", [(128, 256, 512), (256, 256, 54000), (128, 64, 96), (100, 240, 1000), (256, 54000, 128),
                                   (256, 1000, 256), (7, 5, 3)])
def test_tc_gemm_matches_float64(M, N, K):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A, Bm = torch.randn(M, K, generator=g), torch.randn(K, N, generator=g)
    want = A.double() @ Bm.double()
    scale = float(want.abs().max())
    Ad, Bd = A.to(DEV), Bm.to(DEV)
    At, Bt = A.t().contiguous().to(DEV), Bm.t().contiguous().to(DEV)
    for (a, sa), (b, sb) in [((Ad, (K, 1)), (Bt, (1, K))), ((Ad, (K, 1)), (Bd, (N, 1))), ((At, (1, M)), (Bd, (N, 1))),
                             ((At, (1, M)), (Bt, (1, K)))]:
        C = torch.full((M, N), float("nan"), device=DEV)
        ws = ops.tc_gemm(a, sa, b, sb, C, M, N, K)
        torch.cuda.synchronize()
        assert int(ws[:1].view(torch.int32)) == 0, "a bounded mbarrier wait expired inside the kernel"
        err = float((C.cpu().double() - want).abs().max()) / scale
        assert err < 2e-5, (sa, sb, err)      # fp32 accumulation over up to 54,000 terms; budget is 1e-4
    bias = torch.randn(N, generator=g)
    C = torch.empty(M, N, device=DEV)
    ops.tc_gemm(Ad, (K, 1), Bt, (1, K), C, M, N, K, bias=bias.to(DEV), relu=True)
    assert float((C.cpu().double() - torch.relu(want + bias.double())).abs().max()) / scale < 2e-5


def test_plain_tf32_would_not_meet_the_budget():
    """Documents why the split is needed: the same product with operands truncated to TF32 is ~1e-3 off."""
    g = torch.Generator().manual_seed(0)
    A, Bm = torch.randn(128, 4096, generator=g), torch.randn(4096, 256, generator=g)
    trunc = lambda x: (x.view(torch.int32) & -8192).view(torch.float32)
    want = A.double() @ Bm.double()
    tf32 = trunc(A).double() @ trunc(Bm).double()
    assert float((tf32 - want).abs().max() / want.abs().max()) > 1e-4

"""SIMT fp32 GEMM vs tcgen05 3xTF32 GEMM at the encoder's shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
from gennbv_b200 import ops
DEV = "cuda:0"
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for B in (128, 256):
    for name, (M, N, K, amode, bmode) in {"grid_fc fwd": (B, 256, 54000, "k", "k"), "grid_fc dX": (B, 54000, 256, "k", "n"),
                                         "grid_fc dW": (256, 54000, B, "m", "n"), "act_fc1 fwd": (B, 256, 2400, "k", "k"),
                                         "out_fc fwd": (B, 256, 512, "k", "k")}.items():
        A = torch.randn((M, K) if amode == "k" else (K, M), device=DEV)
        Bm = torch.randn((N, K) if bmode == "k" else (K, N), device=DEV)
        sa = (K, 1) if amode == "k" else (1, M)
        sb = (1, K) if bmode == "k" else (N, 1)
        C = torch.empty(M, N, device=DEV)
        ws = torch.zeros(64 + 200 * M * N if M * N < 10_000_000 else 64 + 2 * M * N, device=DEV)
        t_tc = timeit(lambda: ops.tc_gemm(A, sa, Bm, sb, C, M, N, K, workspace=ws))
        t_simt = timeit(lambda: ops.sgemm(A, sa, Bm, sb, C, M, N, K))
        fl = 2.0 * M * N * K
        print(f"B={B} {name:12s} M={M} N={N} K={K}: simt {t_simt:.3f} ms ({fl/t_simt/1e9:.1f} TF)  tc3x {t_tc:.3f} ms ({fl/t_tc/1e9:.1f} TF)")

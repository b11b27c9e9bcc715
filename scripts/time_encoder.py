"""Stage timings of the policy kernels at BASELINE shapes (not a bench line; feeds DESIGN.md / kernel tuning)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from test_policy_gpu import make_policy, DEV

G = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for B in (128, 256):
    pol, ref, D = make_policy(G, 1)
    enc = pol.features_extractor
    obs = torch.zeros(B, D, device=DEV)
    obs[:, :600] = torch.randn(B, 600, device=DEV)
    obs[:, 600:600 + G ** 3] = torch.randint(-1, 2, (B, G ** 3), device=DEV).float()
    pol.train(True)
    dfeat = torch.randn(B, 256, device=DEV)
    grads = pol.encoder_grad_views()
    def fwd():
        return enc._run_forward(obs, need_bwd=True, training=True)
    def bwd(f):
        enc._run_backward(obs, f, dfeat, B, True, enc._ws, grads=grads)
    for _ in range(3):
        f = fwd(); bwd(f)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    n = 10
    tf = tb = 0.0
    for _ in range(n):
        ev[0].record(); f = fwd(); ev[1].record(); bwd(f); ev[2].record()
        torch.cuda.synchronize()
        tf += ev[0].elapsed_time(ev[1]); tb += ev[1].elapsed_time(ev[2])
    pol.train(False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        enc(obs); torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            enc(obs)
        e1.record(); torch.cuda.synchronize()
    print(f"G={G} B={B}: fwd(train) {tf/n:.3f} ms  bwd {tb/n:.3f} ms  fwd(eval) {e0.elapsed_time(e1)/n:.3f} ms")

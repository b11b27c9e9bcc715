"""Diagnostic: per-tensor relative error of the encoder gradients vs the float64 oracle at a given (G, B)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch
from test_policy_gpu import make_policy, oracle_grads, rel_err, DEV, kernel_relu_masks

G, B, seed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 5
pol, ref, D = make_policy(G, seed)
g = torch.Generator().manual_seed(seed)
obs = torch.zeros(B, D)
obs[:, :600] = torch.randn(B, 600, generator=g) * 3
obs[:, 600:600 + G ** 3] = torch.randint(-1, 2, (B, G ** 3), generator=g).float()
wsum = torch.randn(B, 256, generator=g)
enc = pol.features_extractor
for training in (False, True):
    f32, g32, b32 = oracle_grads(ref, obs, wsum, training, torch.float32)
    pol.train(training)
    for p in enc.parameters():
        p.grad = None
    f = enc(obs.to(DEV))
    masks = kernel_relu_masks(enc, B)
    (f * wsum.to(DEV)).sum().backward()
    f64, g64, b64 = oracle_grads(ref, obs, wsum, training, torch.float64, masks=masks)
    fn, gn, bn = oracle_grads(ref, obs, wsum, training, torch.float64)
    print("   tie decisions differing from the float64 oracle's own:", [int((m.double() - 0).sum()) for m in masks], "set; natural-vs-forced max grad diff",
          max(rel_err(gn[k], g64[k]) for k in g64 if not k.endswith("grid.0.bias") and not k.endswith("grid.3.bias")))
    print(f"G={G} B={B} training={training} features err {rel_err(f.detach().cpu(), f64):.2e}")
    for k, p in enc.named_parameters():
        print(f"   {k:40s} vs f64 {rel_err(p.grad.cpu(), g64[k]):.2e}   f32-oracle vs f64 {rel_err(g32[k], g64[k]):.2e}   |g|max {float(g64[k].abs().max()):.3e}")

import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import encoder_ref
from gennbv_b200 import _lib, ops
from test_policy_gpu import make_policy, rel_err, golden_obs, DEV

g = torch.Generator().manual_seed(0)
B = 128
lp = (torch.randn(B, generator=g) * 0.3 - 8).requires_grad_()
ent = (torch.rand(B, generator=g) * 3 + 5).requires_grad_()
v = torch.randn(B, generator=g).requires_grad_()
old_v, old_lp = v.detach() + 0.3 * torch.randn(B, generator=g), lp.detach() + 0.25 * torch.randn(B, generator=g)
adv, ret = torch.randn(B, generator=g) * 2, torch.randn(B, generator=g)
loss, parts = encoder_ref.ppo_loss(v, lp, ent, old_v, old_lp, adv, ret)
loss.backward()
d = lambda t: t.detach().to(DEV).contiguous()
sc = torch.zeros(8, device=DEV)
glp, ge, gv = (torch.empty(B, device=DEV) for _ in range(3))
args = [d(x) for x in (lp, ent, v, old_v, old_lp, adv, ret)]
rc = _lib.lib().gnbv_ppo_loss(*[a.data_ptr() for a in args], B, 0.2, 0.2, 0.01, 0.8, 10.0, 1, sc.data_ptr(),
                              glp.data_ptr(), ge.data_ptr(), gv.data_ptr(), ops._stream())
print("rc", rc, "scalars", sc.cpu().tolist())
print("torch", float(loss), {k: float(x) for k, x in parts.items()}, "adv mean/std", float(adv.mean()), float(adv.std()))
print("glp err", rel_err(glp.cpu(), lp.grad), "ge", rel_err(ge.cpu(), ent.grad), "gv", rel_err(gv.cpu(), v.grad))

for G, Bn, seed in [(20, 9, 1)]:
    pol, ref, D = make_policy(G, seed)
    gg = torch.Generator().manual_seed(seed)
    obs = torch.zeros(Bn, D)
    obs[:, :600] = torch.randn(Bn, 600, generator=gg) * 3
    obs[:, 600:600 + G ** 3] = torch.randint(-1, 2, (Bn, G ** 3), generator=gg).float()
    wsum = torch.randn(Bn, 256, generator=gg)
    for training in (False, True):
        ref.train(training); pol.train(training)
        ref.zero_grad()
        f_ref = ref.features_extractor(obs)
        (f_ref * wsum).sum().backward()
        enc = pol.features_extractor
        for p in enc.parameters():
            p.grad = None
        f = enc(obs.to(DEV))
        (f * wsum.to(DEV)).sum().backward()
        rg = dict(ref.features_extractor.named_parameters())
        print("training", training, "feat err", rel_err(f.detach().cpu(), f_ref.detach()))
        for k, p in enc.named_parameters():
            a, b = p.grad.cpu().double(), rg[k].grad.double()
            print(f"   {k:40s} rel {rel_err(a, b):.3e}  |ref|max {float(b.abs().max()):.3e} abs err {float((a-b).abs().max()):.3e}")
    # heads
    actions = torch.stack([torch.randint(0, n, (Bn,), generator=gg) for n in (81, 81, 51, 1, 13, 13)], 1)
    ref.zero_grad()
    for p in pol.parameters():
        p.grad = None
    w3 = torch.randn(3, Bn, generator=gg)
    vr, lr_, er = ref.evaluate_actions(obs, actions)
    ((vr.flatten() * w3[0]).sum() + (lr_ * w3[1]).sum() + (er * w3[2]).sum()).backward()
    vm, lm, em = pol.evaluate_actions(obs.to(DEV), actions.to(DEV))
    w3d = w3.to(DEV)
    ((vm.flatten() * w3d[0]).sum() + (lm * w3d[1]).sum() + (em * w3d[2]).sum()).backward()
    print("heads fwd", rel_err(vm.detach().cpu(), vr.detach()), rel_err(lm.detach().cpu(), lr_.detach()), rel_err(em.detach().cpu(), er.detach()))
    rg = dict(ref.named_parameters())
    for k, p in pol.named_parameters():
        a, b = p.grad.cpu().double(), rg[k].grad.double()
        print(f"   {k:55s} rel {rel_err(a, b):.3e}  |ref|max {float(b.abs().max()):.3e}")

import ctypes, os, sys, json, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch
from test_policy_gpu import make_policy, DEV
from gennbv_b200 import _lib
G, B = 64, 256
pol, ref, D = make_policy(G, 1)
obs = torch.zeros(B, D, device=DEV)
obs[:, 600:600 + G ** 3] = torch.randint(-1, 2, (B, G ** 3), device=DEV).float()
pol.train(True)
enc = pol.features_extractor
for _ in range(3):
    with torch.no_grad(): enc(obs)
torch.cuda.synchronize()
buf = (ctypes.c_uint64 * 32)()
_lib.check(_lib.lib().gnbv_debug_ts_profile(buf), "gnbv_debug_ts_profile")
v = list(buf)
names = {0: "loader.wait_pfree", 1: "loader.wait_raw", 2: "loader.permute", 4: "prod.wait_pfull", 5: "prod.wait_free", 6: "prod.work",
         8: "mma.wait_accfree", 9: "mma.wait_full", 10: "mma.issue", 12: "epi.wait_accfull", 13: "epi.work", 14: "epi.total", 15: "tiles"}
tiles = max(1, v[15])
for k, n in names.items():
    print(f"{n:20s} {v[k]:12d}  per tile {v[k] / tiles:10.1f}")

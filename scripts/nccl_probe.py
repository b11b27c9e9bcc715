"""Times the flat-gradient all-reduce alone and interleaved with compute (2+ ranks, torchrun)."""
import os, time, torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 14_700_000
g = torch.randn(n, device="cuda")
a = torch.randn(8192, 8192, device="cuda")
def timeit(fn, k=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
t_ar = timeit(lambda: dist.all_reduce(g))
t_mm = timeit(lambda: a @ a)
t_both = timeit(lambda: (a @ a, dist.all_reduce(g)))
if dist.get_rank() == 0:
    print(f"all_reduce {n*4/1e6:.0f} MB: {t_ar:.3f} ms ({n*4/t_ar/1e6:.0f} GB/s algbw); matmul {t_mm:.3f} ms; interleaved {t_both:.3f} ms")
dist.destroy_process_group()

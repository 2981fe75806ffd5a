"""times the pieces of sharding.gather_surface_to_root on real stripe surfaces (16384^2 over the ranks)"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
import vkvg_b200 as v
from vkvg_b200 import sharding
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
os.environ.setdefault("VKVG_B200_DEVICE", str(local))
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
H = W = 16384
dev = v.Device(4)
surf, y0, h = sharding.stripe_surface(dev, W, H, rank, world)
out = None
res = {}
def timed(name, fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    res[name] = (time.perf_counter() - t0) / n * 1e3
timed("as_tensor", lambda: surf.as_tensor())
t = surf.as_tensor()
res["ptr_equal"] = int(t.data_ptr()) == int(v.lib().vkvg_b200_surface_device_pointer(surf.h))
def g():
    global out
    out = sharding.gather_surface_to_root(surf, H, out=out)
timed("gather_surface_to_root", g)
def g2():
    global out
    out = sharding.gather_to_root(t[:h], H, out=out)
timed("gather_to_root(view)", g2)
if rank == 0:
    print(json.dumps(res))
dist.destroy_process_group()

"""Which collective moves the stripes of a 16384^2 RGBA8 surface to rank 0 fastest: NCCL send / recv straight into the rows of the root's
buffer (sharding.gather_to_root), or an all-gather into every rank's buffer (sharding.gather_stripes)?
    torchrun --nproc-per-node N profiles/scripts/gather_bench.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
from vkvg_b200 import sharding
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
H = W = 16384
y0, h = sharding.stripe_rows(H, world)[rank]
stripe = torch.full((h, W, 4), rank + 1, dtype=torch.uint8, device="cuda")
out = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda") if rank == 0 else None
res = {}
for name, fn in (("p2p_to_root", lambda: sharding.gather_to_root(stripe, H, out=out)), ("all_gather", lambda: sharding.gather_stripes(stripe, H))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 5], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[name] = float(t.item())
if rank == 0:
    res["world"] = world; res["GiB"] = H * W * 4 / 2**30
    res["env"] = {k: v for k, v in os.environ.items() if k.startswith("NCCL")}
    print(json.dumps(res))
dist.destroy_process_group()

"""GPU: a stripe delivered into another surface's memory by the flush itself (vkvg_b200_surface_set_readback with a DEVICE address:
what sharding.deliver_to_root sets up across GPUs through CUDA IPC).  One GPU: the target is a second surface of the same device.
Two or more GPUs: two processes, the root's picture opened over IPC, NVLink copies - skipped on a one-GPU box."""
import os
import subprocess
import sys

import numpy as np
import pytest

import vkvg_b200 as v
from tests import scenes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _emit(c, W, H, n=300):
    polys, cols = scenes.polygons_c2(n, W, 11)
    for i, (p, col) in enumerate(zip(polys, cols)):
        p = p * np.array([1.0, H / W], np.float32)
        c.set_source_rgba(*[float(x) for x in col])
        c.set_fill_rule(i % 2)
        c.move_to(float(p[0, 0]), float(p[0, 1]))
        for q in p[1:]:
            c.line_to(float(q[0]), float(q[1]))
        c.close_path()
        c.fill()


@pytest.mark.parametrize("W,H,world", [(640, 1000, 3), (2048, 4096, 2)])   # one band per stripe | four bands per stripe
def test_stripes_delivered_into_a_full_surface_on_the_same_device(dev4, W, H, world):
    from vkvg_b200 import sharding
    whole = v.Surface(dev4, W, H)
    c = v.Context(whole)
    _emit(c, W, H)
    c.flush()
    want = whole.pixels()
    full = v.Surface(dev4, W, H)          # the "root's picture": only ever written by the deliveries
    base = full.device_pointer()
    for rank in range(world):
        surf, y0, h = sharding.stripe_surface(dev4, W, H, rank, world)
        surf.set_readback(base + y0 * W * 4)
        ctx = v.Context(surf)
        for _ in range(3):                 # the third flush of one structure replays a captured graph: the copies are part of it
            ctx.clear()
            _emit(ctx, W, H)
            ctx.flush()
        surf.wait_delivered(base + y0 * W * 4)
        ctx.close()
        surf.set_readback(None)
        surf.close()
    dev4.synchronize()
    assert np.array_equal(full.pixels(), want)


_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
import vkvg_b200 as v
from vkvg_b200 import sharding
from tests.test_gpu_delivery import _emit
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
os.environ["VKVG_B200_DEVICE"] = str(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
W, H = 2048, 4096
dev = v.Device(4)
surf, y0, h = sharding.stripe_surface(dev, W, H, rank, world)
peer = sharding.deliver_to_root(surf, y0, h, W, H)
ctx = v.Context(surf)
for _ in range(3):
    ctx.clear(); _emit(ctx, W, H); ctx.flush()
    peer.barrier()
if rank == 0:
    whole = v.Surface(dev, W, H)
    c = v.Context(whole); _emit(c, W, H); c.flush()
    assert np.array_equal(peer.full.pixels(), whole.pixels()), "the root's picture differs from the unsharded render"
    print("DELIVERY_OK")
dist.barrier()
peer.close()
dist.destroy_process_group()
'''


def test_stripes_delivered_to_the_root_gpu_over_ipc(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29731",
                        str(script)], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0 and "DELIVERY_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]

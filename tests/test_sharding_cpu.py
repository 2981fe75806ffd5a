"""CPU, world_size 2 over gloo: the multi-GPU host logic (stripe partition, ragged all-gather, canvas round-robin).
Rendering is stubbed with the oracle (the CUDA library cannot run here); what is tested is the plumbing of
vkvg_b200/sharding.py, which the GPU path uses unchanged with the NCCL backend."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vkvg_b200 import sharding

W, H = 96, 88  # 88 rows = 5.5 tiles: ragged last stripe


def _scene(g):
    rng = np.random.default_rng(3)
    g.set_fill_rule(0)
    for _ in range(12):
        k = int(rng.integers(3, 8))
        pts = np.round(rng.uniform(2, [W - 2, H - 2], (k, 2)) * 8) / 8  # 1/8 px grid: translation by whole tiles is exact
        g.set_source_rgba(*[float(x) for x in rng.uniform(0.2, 1.0, 4)])
        g.move_to(float(pts[0, 0]), float(pts[0, 1]))
        for p in pts[1:]:
            g.line_to(float(p[0]), float(p[1]))
        g.close_path()
        g.fill()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from oracle import Oracle
    y0, h = sharding.stripe_rows(H, world)[rank]
    o = Oracle(W, max(h, 1), 4)
    o.translate(0.0, -float(y0))
    _scene(o)
    local = torch.from_numpy(o.pixels()[:h].copy())
    full = sharding.gather_stripes(local, H).numpy()
    ref = Oracle(W, H, 4)
    _scene(ref)
    assert np.array_equal(full, ref.pixels()), "rank %d: gathered stripes differ from the unsharded render" % rank
    # gather to one root: every stripe straight into its rows of the root's buffer, nothing on the other ranks
    root_img = sharding.gather_to_root(local, H, root=0)
    if rank == 0:
        assert np.array_equal(root_img.numpy(), ref.pixels())
    else:
        assert root_img is None
    # independent canvases: every canvas is rendered by exactly one rank
    mine = torch.zeros(37, dtype=torch.int64)
    mine[sharding.canvases_for_rank(37, rank, world)] = 1
    dist.all_reduce(mine)
    assert bool((mine == 1).all())
    dist.destroy_process_group()


def test_stripe_rows_partition():
    for height in (16, 17, 88, 1024, 4096, 16384, 5):
        for world in (1, 2, 3, 4, 8):
            rows = sharding.stripe_rows(height, world)
            assert len(rows) == world
            assert rows[0][0] == 0 and sum(h for _, h in rows) == height
            y = 0
            for y0, h in rows:
                assert y0 == y and y0 % 16 == 0 and h >= 0
                y += h
    assert sharding.stripe_rows(16384, 8) == [(2048 * r, 2048) for r in range(8)]


def test_canvas_round_robin():
    got = sorted(sum((sharding.canvases_for_rank(1024, r, 8) for r in range(8)), []))
    assert got == list(range(1024))
    assert all(len(sharding.canvases_for_rank(1024, r, 8)) == 128 for r in range(8))


@pytest.mark.timeout(180)
def test_striped_render_and_gather_world2(oracle_lib):
    mp.spawn(_worker, args=(2, _free_port()), nprocs=2, join=True)


def _worker_more_ranks_than_rows(rank, world, port):
    """a 20-row surface has two tile rows: the third rank owns no rows, must still take part in both gathers and not hang"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    Hs = 20
    rows = sharding.stripe_rows(Hs, world)
    assert any(h == 0 for _, h in rows)
    y0, h = rows[rank]
    full_ref = (np.arange(Hs * 8 * 4, dtype=np.uint32) % 251).astype(np.uint8).reshape(Hs, 8, 4)
    local = torch.from_numpy(full_ref[y0:y0 + h].copy()) if h else torch.zeros((0, 8, 4), dtype=torch.uint8)
    assert np.array_equal(sharding.gather_stripes(local, Hs).numpy(), full_ref)
    img = sharding.gather_to_root(local, Hs, root=0)
    assert (rank == 0 and np.array_equal(img.numpy(), full_ref)) or (rank != 0 and img is None)
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_gathers_with_more_ranks_than_tile_rows():
    mp.spawn(_worker_more_ranks_than_rows, args=(3, _free_port()), nprocs=3, join=True)

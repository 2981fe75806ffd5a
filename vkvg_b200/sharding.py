"""Multi-GPU sharding of the path-rendering hot path.  One process per GPU; nothing on the data path
needs an exchange:

  * tile-row stripes of one large surface: rank r renders rows [y0_r, y0_r + h_r) of the logical surface by issuing the
    SAME drawing calls on a stripe surface (`vkvg_b200_surface_create_stripe`); the rows are gathered afterwards with one
    all-gather of contiguous RGBA8 rows (NCCL over NVLink on GPUs, gloo in the CPU tests) only when an assembled image is
    needed;
  * independent canvases: canvas i belongs to rank i mod world; no collective at all.

The functions here are pure host logic plus `torch.distributed` plumbing; rendering itself is done by whoever calls them
(the CUDA library on GPUs; the CPU tests plug in the oracle to check the plumbing).
"""
TILE = 16


def stripe_rows(height, world, tile=TILE):
    """Contiguous blocks of tile rows per rank: [(y0, h)] * world; ranks past the last tile row get (height, 0)."""
    rows = (height + tile - 1) // tile
    out = []
    for r in range(world):
        a = rows * r // world
        b = rows * (r + 1) // world
        y0, y1 = min(a * tile, height), min(b * tile, height)
        out.append((y0, y1 - y0))
    return out


def canvases_for_rank(n_canvases, rank, world):
    """round-robin assignment of independent canvases"""
    return list(range(rank, n_canvases, world))


def gather_stripes(local, height, group=None):
    """all-gather stripes of shape (h_r, W, 4) uint8 (torch tensors on the backend's device) into the (height, W, 4) image.
    Stripes may differ in height (ragged last block): every rank pads to the tallest, gathers, and crops."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rows = stripe_rows(height, world)
    hmax = max(h for _, h in rows)
    w = local.shape[1]
    pad = torch.zeros((hmax, w, 4), dtype=torch.uint8, device=local.device)
    pad[:local.shape[0]] = local
    out = torch.empty((world, hmax, w, 4), dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(out.view(-1), pad.view(-1), group=group)
    full = torch.empty((height, w, 4), dtype=torch.uint8, device=local.device)
    for r, (y0, h) in enumerate(rows):
        full[y0:y0 + h] = out[r, :h]
    return full


def render_striped(dev, width, height, emit, rank, world):
    """GPU path: render this rank's stripe of a width x height surface; returns (Surface, y0, h).  `emit(ctx)` issues the
    drawing calls of the WHOLE scene (every rank replays all of it; geometry outside the stripe is culled by the tile
    binning)."""
    import vkvg_b200 as v
    y0, h = stripe_rows(height, world)[rank]
    surf = v.Surface(dev, width, max(h, 1), full_height=height, origin_y=y0)
    ctx = v.Context(surf)
    emit(ctx)
    ctx.flush()
    ctx.close()
    return surf, y0, h


def gather_surface(surf, y0, h, height, group=None):
    """GPU path: NCCL all-gather of the stripes rendered by render_striped -> (height, W, 4) uint8 CUDA tensor."""
    import torch
    local = torch.empty((max(h, 1), surf.width, 4), dtype=torch.uint8, device="cuda")
    surf.copy_to_device(local.data_ptr())
    torch.cuda.synchronize()
    return gather_stripes(local[:h], height, group)

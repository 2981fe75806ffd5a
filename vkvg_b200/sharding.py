"""Multi-GPU sharding of the path-rendering hot path.  One process per GPU; nothing on the data path
needs an exchange:

  * tile-row stripes of one large surface: rank r renders rows [y0_r, y0_r + h_r) of the logical surface by issuing the
    SAME drawing calls on a stripe surface (`vkvg_b200_surface_create_stripe`); the rows are gathered afterwards with one
    all-gather of contiguous RGBA8 rows (NCCL over NVLink on GPUs, gloo in the CPU tests) only when an assembled image is
    needed;
    or - `deliver_to_root` - every rank names its rows of the ROOT's surface (opened through CUDA IPC) as the read-back target of
    its stripe: finished bands of tile rows cross NVLink on the copy engines while later bands still render, and nothing follows
    the frame but a barrier;
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
    """all-gather stripes of shape (h_r, W, 4) uint8 (torch tensors on the backend's device) into the (height, W, 4) image on every
    rank.  Equal heights (the usual case: 16384 rows over 2 / 4 / 8 ranks): one all_gather_into_tensor straight into the final buffer,
    no staging.  Ragged heights or ranks without rows: every rank pads to the tallest stripe, gathers, and crops."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rows = stripe_rows(height, world)
    w = local.shape[1]
    hs = [h for _, h in rows]
    if min(hs) == max(hs) and sum(hs) == height:
        full = torch.empty((height, w, 4), dtype=torch.uint8, device=local.device)
        dist.all_gather_into_tensor(full.view(-1), local[:hs[0]].contiguous().view(-1), group=group)
        return full
    hmax = max(hs)
    pad = torch.zeros((hmax, w, 4), dtype=torch.uint8, device=local.device)
    n = min(local.shape[0], rows[dist.get_rank(group)][1])
    pad[:n] = local[:n]
    out = torch.empty((world, hmax, w, 4), dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(out.view(-1), pad.view(-1), group=group)
    full = torch.empty((height, w, 4), dtype=torch.uint8, device=local.device)
    for r, (y0, h) in enumerate(rows):
        full[y0:y0 + h] = out[r, :h]
    return full


def gather_to_root(local, height, out=None, root=0, group=None):
    """Stripes (h_r, W, 4) uint8 -> the (height, W, 4) image on `root` (None elsewhere).  Every stripe travels once, straight into its
    rows of the final buffer: the root posts one receive per sender into a view of `out`, senders send from `local` (which may be the
    surface's own memory, Surface.as_tensor(): no staging copy on either side).  Ragged heights and ranks without rows (more ranks
    than tile rows) are fine: such a rank sends nothing."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    rows = stripe_rows(height, world)
    w = local.shape[1]
    ops = []
    if rank == root:
        if out is None:
            out = torch.empty((height, w, 4), dtype=torch.uint8, device=local.device)
        for r, (y0, h) in enumerate(rows):
            if h == 0:
                continue
            if r == root:
                out[y0:y0 + h].copy_(local[:h])
            else:
                ops.append(dist.P2POp(dist.irecv, out[y0:y0 + h], r, group))
    elif rows[rank][1] > 0:
        ops.append(dist.P2POp(dist.isend, local[:rows[rank][1]], root, group))
    if ops:
        for q in dist.batch_isend_irecv(ops):
            q.wait()
    return out if rank == root else None


def stripe_surface(dev, width, height, rank, world):
    """the stripe surface of `rank` (one tile row at the bottom of the logical surface for a rank without rows, so that every rank can
    issue the same calls); returns (Surface, y0, h) with h the rows that count."""
    import vkvg_b200 as v
    y0, h = stripe_rows(height, world)[rank]
    if h == 0:
        last = ((height - 1) // TILE) * TILE
        return v.Surface(dev, width, height - last, full_height=height, origin_y=last), height, 0
    return v.Surface(dev, width, h, full_height=height, origin_y=y0), y0, h


def render_striped(dev, width, height, emit, rank, world):
    """GPU path: render this rank's stripe of a width x height surface; returns (Surface, y0, h).  `emit(ctx)` issues the
    drawing calls of the WHOLE scene (every rank replays all of it; geometry outside the stripe is culled by the tile
    binning)."""
    import vkvg_b200 as v
    surf, y0, h = stripe_surface(dev, width, height, rank, world)
    ctx = v.Context(surf)
    emit(ctx)
    ctx.flush()
    ctx.close()
    return surf, y0, h


def gather_surface(surf, y0, h, height, group=None):
    """GPU path: NCCL all-gather of the stripes rendered by render_striped -> (height, W, 4) uint8 CUDA tensor on every rank."""
    surf.dev.synchronize()
    return gather_stripes(surf.as_tensor()[:h], height, group)


def gather_surface_to_root(surf, height, out=None, root=0, group=None):
    """GPU path: the stripes rendered on stripe surfaces gathered into one image on `root`, sent straight from the surfaces' memory."""
    import torch.distributed as dist
    surf.dev.synchronize()   # the library renders on its own stream
    h = stripe_rows(height, dist.get_world_size(group))[dist.get_rank(group)][1]
    return gather_to_root(surf.as_tensor()[:max(h, 0)], height, out=out, root=root, group=group)


class StripeDelivery:
    """What deliver_to_root returns.  `full` is the root's whole-picture Surface (None elsewhere); `target` the address this rank's stripe
    is delivered to (0 for a rank without rows).  wait() returns when THIS rank's rows of the last flush have arrived; the ranks then
    meet at whatever barrier they share (dist.barrier) before the root reads `full`."""

    def __init__(self, surf, full, target, base, group):
        self.surf, self.full, self.target, self._base, self.group = surf, full, target, base, group

    def wait(self):
        if self.target:
            self.surf.wait_delivered(self.target)
        else:
            self.surf.dev.synchronize()

    def barrier(self):
        import torch.distributed as dist
        self.wait()
        dist.barrier(self.group)

    def close(self):
        self.surf.set_readback(None)
        if self._base:
            self.surf.dev.ipc_close(self._base)
            self._base = 0


def deliver_to_root(surf, y0, h, width, height, root=0, group=None):
    """GPU path, no collective on the data path: the root creates the width x height surface and exports its image (cudaIpcMemHandle_t),
    every other rank opens it and sets rows [y0, y0 + h) of it as the read-back target of its stripe surface `surf` (the root does the same
    with a plain device pointer).  From then on every flush of `surf` lands in the root's picture band by band, overlapped with the
    rendering of the later bands (vkvg_b200_surface_set_readback).  Returns a StripeDelivery."""
    import vkvg_b200 as v
    import torch.distributed as dist
    rank = dist.get_rank(group)
    full = v.Surface(surf.dev, width, height) if rank == root else None
    box = [full.ipc_export() if rank == root else None]
    dist.broadcast_object_list(box, src=root, group=group)
    base = 0
    if rank == root:
        addr = full.device_pointer()
    else:
        base = addr = surf.dev.ipc_open(box[0])
    target = addr + y0 * width * 4 if h > 0 else 0
    surf.set_readback(target or None)
    return StripeDelivery(surf, full, target, base, group)

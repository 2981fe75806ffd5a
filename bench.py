#!/usr/bin/env python
"""bench.py — the driver's benchmark contract for the vkvg path-rendering hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c3|c4|c5a|c5b|blit] [--rule nz|eo] [--coverage msaa|analytic]
                    [--no-graph] [--impl reference]

One "step" = one pass of the hot path (flatten -> stroke/fill edges -> tile binning -> winding + paint + OVER) over one
synthetic scene.  Default workload = BASELINE.json configs[1] ("c2": 100k random self-intersecting polygons, 4096x4096,
4 samples).  At N > 1 every rank renders its own independent canvas of the same configuration (weak scaling, no
collective on the data path: SURVEY.md §8e "independent canvases").

  value  = whole-job Mpix/s with the recorded scene already resident in HBM (device-timed, CUDA events on the library's
           stream, max over ranks, L2 flushed between steps outside the timed events); each step is one replay of the CUDA graph the
           library captures for a repeating frame (--no-graph: plain launches); a separate pass with plain launches and events
           between the pipeline stages supplies stage_ms
  e2e    = same metric through the public C ABI with HOST buffers: command arrays in pinned host memory -> vkvg_b200_replay
           -> vkvg_flush -> read the surface back to pinned host memory, all inside the timed region (wall clock, max over ranks)
  --impl reference: the reference's own CPU implementation of the path (oracle/_ref = its unmodified tessellation sources,
           plus the oracle's scalar restatement of the Vulkan rasteriser it delegates to) on the host cores, bounded sample.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SIZES = {"c1": 1024, "c2": 4096, "c3": 4096, "c4": 8192, "c5a": 16384, "c5b": 1024, "blit": 4096}
C5B_CANVASES, C5B_BATCH = 1024, 16   # 1024 independent 1024^2 canvases, 16 per flush (one batch surface of 1024 x 16384)
UNITS = {"c1": ("tiger_frames_per_s", "frames/s"), "c2": ("fill_Mpix_per_s", "Mpix/s"), "c3": ("stroke_Msegments_per_s", "Msegments/s"),
         "c4": ("fill_Mpix_per_s", "Mpix/s"), "c5a": ("fill_Mpix_per_s", "Mpix/s"), "c5b": ("tiger_frames_per_s", "frames/s"),
         "blit": ("fill_Mpix_per_s", "Mpix/s")}
WORKLOAD_NAMES = {
    "c1": "C1 tiger.svg via nanoSVG, 1024x1024, 4 samples, even-odd fills + miter strokes",
    "c2": "C2 100k random self-intersecting polygons, one fill each, 4096x4096, 4 samples",
    "c3": "C3 1M-segment polyline stroke, width 3, round joins/caps, dash {10,6}, 4096x4096, 4 samples",
    "c4": "C4 50k cubic-Bezier paths, linear/radial gradient fills, OVER, 8192x8192, 4 samples",
    "c5b": "C5b batch of 1024 independent 1024x1024 tiger canvases (per-canvas affine jitter), 16 canvases per flush in one batch surface, canvases split across ranks",
    "blit": "layer compositing: 8 translucent 2048x2048 surface sources (bilinear, rotated) painted over a 4096x4096 surface (SURVEY §8f rank 2)",
    "c5a": "C5a 16384x16384 surface, 5M-segment mix (C2-style polygons + C3-style dashed polylines), sharded by tile-row stripes, NCCL all-gather",
}


# ---------------------------------------------------------------------------------------------------------------
# scenes -> packed command stream (ops, args); `emit` drives any object with the drawing vocabulary
# ---------------------------------------------------------------------------------------------------------------
def build_scene(workload, seed, rule, n_limit=None, first=0):
    """returns (emit(g), units) where emit replays the scene on g and units is the metric's unit count of one step."""
    from tests import scenes
    size = SIZES[workload]
    if workload == "c2":
        polys, cols = scenes.polygons_c2(100000, size, seed)
        sl = slice(first, first + n_limit) if n_limit else slice(None)
        polys, cols = polys[sl], cols[sl]

        def emit(g):
            g.set_fill_rule(0 if rule == "eo" else 1)
            for p, c in zip(polys, cols):
                g.set_source_rgba(*[float(x) for x in c])
                poly(g, p)
                g.close_path()
                g.fill()
        return emit, size * size / 1e6, dict(n_paths=len(polys), n_segments=int(sum(len(p) for p in polys)))
    if workload == "c3":
        n = 1_000_001 if not n_limit else n_limit + 1
        pts = scenes.polyline_c3(n, size, seed)

        def emit(g):
            g.set_source_rgba(0.1, 0.2, 0.8, 1.0)
            g.set_line_width(3.0)
            g.set_line_join(1)
            g.set_line_cap(1)
            g.set_dash([10.0, 6.0], 0.0)
            poly(g, pts)
            g.stroke()
        return emit, (n - 1) / 1e6, dict(n_paths=1, n_segments=n - 1)
    if workload == "c4":
        paths = scenes.curves_c4(50000, size, seed)
        sl = slice(first, first + n_limit) if n_limit else slice(None)
        paths = paths[sl]

        def emit(g):
            g.set_fill_rule(1)
            for pts, kind, (cx, cy), stops in paths:
                if kind == 0:
                    g.set_source_linear(cx - 64, cy - 64, cx + 64, cy + 64, stops)
                else:
                    g.set_source_radial(cx, cy, 4.0, cx + 8, cy - 8, 96.0, stops)
                g.move_to(float(pts[-1, 2, 0]), float(pts[-1, 2, 1]))
                for s in pts:
                    g.curve_to(*[float(x) for x in s.ravel()])
                g.close_path()
                g.fill()
        return emit, size * size / 1e6, dict(n_paths=len(paths), n_segments=int(sum(len(p[0]) for p in paths)))
    if workload == "c5a":
        n_poly = 238000 if not n_limit else n_limit
        polys, cols = scenes.polygons_c2(n_poly, size, seed)
        # polygons_c2 draws radii for a 4096 canvas; keep them (small shapes on a huge surface), plus 25 polylines of 100k segments
        lines = [scenes.polyline_c3(100001, size, seed * 100 + i) for i in range(25 if not n_limit else 1)]

        def emit(g):
            g.set_fill_rule(1)
            for p, c in zip(polys, cols):
                g.set_source_rgba(*[float(x) for x in c])
                poly(g, p)
                g.close_path()
                g.fill()
            g.set_line_width(3.0)
            g.set_line_join(1)
            g.set_line_cap(1)
            g.set_dash([10.0, 6.0], 0.0)
            for i, pts in enumerate(lines):
                g.set_source_rgba(0.1 + 0.03 * i, 0.2, 0.8 - 0.02 * i, 1.0)
                poly(g, pts)
                g.stroke()
        nseg = int(sum(len(p) for p in polys)) + sum(len(l) - 1 for l in lines)
        return emit, size * size / 1e6, dict(n_paths=len(polys) + len(lines), n_segments=nseg)
    if workload == "blit":
        n_layers = n_limit or 8

        def emit(g):
            g.set_source_rgba(0.1, 0.1, 0.12, 1.0)
            g.paint()
            for k in range(n_layers):
                g.identity_matrix()
                g.translate(300.0 * k + 100.0, 180.0 * k + 60.0)
                g.rotate(0.11 * k)
                g.set_opacity(0.6 + 0.05 * k)
                g.set_source_layer(4 if k % 2 else 3)   # bilinear / nearest
                g.rectangle(0.0, 0.0, 2048.0, 2048.0)
                g.fill()
        return emit, size * size / 1e6, dict(n_paths=n_layers + 1, n_segments=4 * n_layers)
    if workload == "c5b":   # one flush worth of canvases; the step replays it for every batch this rank owns
        w, h, shapes = scenes.load_nsvg(os.path.join(ROOT, "tests", "golden", "tiger.nsvg.bin"))
        r = scenes.SplitMix64(900 + seed)
        jit = [(r.uniform(-2, 2), r.uniform(-2, 2)) for _ in range(C5B_BATCH)]

        def emit(g):
            for i, (jx, jy) in enumerate(jit):
                if hasattr(g, "set_canvas"):
                    g.set_canvas(i)
                g.identity_matrix()
                g.translate(jx, jy)
                scenes.render_nsvg(g, shapes)
        return emit, float(C5B_BATCH), dict(n_paths=len(shapes) * C5B_BATCH, n_segments=int(sum((len(p) - 1) // 3 for s in shapes for p, _ in s["paths"])) * C5B_BATCH)
    if workload == "c1":
        w, h, shapes = scenes.load_nsvg(os.path.join(ROOT, "tests", "golden", "tiger.nsvg.bin"))
        if n_limit:
            shapes = shapes[first:first + n_limit]

        def emit(g):
            scenes.render_nsvg(g, shapes)
        return emit, 1.0, dict(n_paths=len(shapes), n_segments=int(sum((len(p) - 1) // 3 for s in shapes for p, _ in s["paths"])))
    raise SystemExit("unknown workload " + workload)


class LayerSource:
    """adds set_source_layer(filter) to a drawing object: the 2048x2048 layer of the blit workload as a surface paint"""

    def __init__(self, g, source):
        self._g, self._source = g, source

    def __getattr__(self, name):
        return getattr(self._g, name)

    def set_source_layer(self, filt):
        self._g.set_source_surface(self._source, 0.0, 0.0, extend=0, filter=filt)


def layer_image():
    from tests.golden import make_golden2 as mg2
    return mg2.checker(2048, 2048, 5)


def poly(g, pts):
    if hasattr(g, "polyline"):
        g.polyline(pts)
    else:
        g.move_to(float(pts[0, 0]), float(pts[0, 1]))
        for p in pts[1:]:
            g.line_to(float(p[0]), float(p[1]))


# ---------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.sm_max = index, False, [], set(), None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.sm.append(int(out[0]))
                self.sm_max = int(out[1])
                bits = int(out[2].strip(), 16)
                names = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
                         0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
                for b, n in names.items():
                    if bits & b and n != "gpu_idle":
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        self.join(timeout=6)
        return {"sm_mhz": int(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons)}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            for k in ("hbm_gbs", "hbm_gbps", "hbm_GBps", "copy_GBps"):
                if k in d:
                    return float(d[k]), "MEASURED_PEAKS.json:" + k
            for k, val in d.items():
                if "hbm" in k.lower() and isinstance(val, (int, float)):
                    return float(val), "MEASURED_PEAKS.json:" + k
                if isinstance(val, dict):
                    for k2, v2 in val.items():
                        if "hbm" in (k + k2).lower() and isinstance(v2, (int, float)):
                            return float(v2), "MEASURED_PEAKS.json:%s.%s" % (k, k2)
        except Exception:
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md); MEASURED_PEAKS.json absent"


# ---------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: reference tessellation object code + oracle raster, on the host cores
# ---------------------------------------------------------------------------------------------------------------
def _ref_worker(args):
    workload, seed, rule, n_limit, first = args
    import oracle
    from oracle import Oracle, Ref
    size = SIZES[workload]
    emit, units, info = build_scene(workload, seed, rule, n_limit=n_limit, first=first)
    kind = "reference" if oracle.ref_available() else "port"
    if workload == "blit":   # textures do not travel through the recorded draw list of the reference build: oracle port only
        kind = "port"
        inner = emit
        img = layer_image()
        emit = lambda g: inner(LayerSource(g, img))  # noqa: E731
    t0 = time.perf_counter()
    o = Oracle(size, size, 4)
    if kind == "reference":
        r = Ref(size, size, 4)
        emit(r)
        r.render_with(o)   # the reference's recorded draw list, rasterised by the Vulkan restatement
        r.close()
    else:
        emit(o)
    o.pixels()
    dt = time.perf_counter() - t0
    o.close()
    return dt, info, kind


SAMPLE = {"blit": 1, "c5b": None, "c1": None, "c2": 1500, "c3": 40000, "c4": 200, "c5a": 1000}  # units of work per host thread per step (paths / segments)
# the single-thread cpu_baseline leg of the default run works on a larger slice: about 10-30 s of CPU work
CPU_SAMPLE = dict(SAMPLE, c2=None)   # the whole C2 scene: ~10 s on the GPU box's host, ~25 s on a slow core
FULL = {"blit": 8, "c5b": 239 * C5B_CANVASES, "c1": 239, "c2": 100000, "c3": 1000000, "c4": 50000, "c5a": 476000}


def reference_step(workload, seed, rule, cores, pool):
    """one bounded step on `cores` host processes; returns (seconds, fraction of the full scene processed, kind)."""
    n = SAMPLE[workload]
    if workload == "c5b":  # independent canvases: one whole tiger canvas per host process
        res = pool.map(_ref_worker, [("c1", seed, rule, None, 0) for _ in range(cores)])
        return max(r[0] for r in res), cores / float(C5B_CANVASES), res[0][2]
    if workload == "c3":   # one polyline: a single context is strictly serial in the reference; threads get separate lines
        jobs = [(workload, seed + i, rule, n, 0) for i in range(cores)]
    elif n is None:
        jobs = [(workload, seed, rule, None, 0) for _ in range(cores)]
    else:
        jobs = [(workload, seed, rule, n, i * n) for i in range(cores)]
    res = pool.map(_ref_worker, jobs)
    dt = max(r[0] for r in res)   # the workers time only the rendering, not the synthetic-scene generation
    frac = cores * (1.0 if n is None else n / FULL[workload])
    return dt, frac, res[0][2]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    import oracle
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        oracle.build(ref=False)
    cores = os.cpu_count() or 1
    w = args.workload
    units_full = build_units(w)
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(min(args.warmup, 1)):
            reference_step(w, 1, args.rule, cores, pool)
        ts, frac, kind = [], 0, "port"
        for _ in range(args.steps):
            dt, frac, kind = reference_step(w, 1, args.rule, cores, pool)
            ts.append(dt)
    t = float(np.mean(ts))
    value = units_full * frac / t
    name, unit = UNITS[w]
    sample = "%d host processes x %s of the %s scene per step (reference tessellation object code + scalar Vulkan-raster restatement)" % (
        cores, "the whole scene" if SAMPLE[w] is None else "%d of %d units" % (SAMPLE[w], FULL[w]), w.upper())
    line = {"impl": "reference", "metric": name, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+i64", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAMES[w], "rule": args.rule, "samples": 4},
            "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def build_units(workload):
    return {"blit": 4096 * 4096 / 1e6, "c5b": float(C5B_CANVASES), "c1": 1.0, "c2": 4096 * 4096 / 1e6, "c3": 1.0, "c4": 8192 * 8192 / 1e6, "c5a": 16384 * 16384 / 1e6}[workload]


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import vkvg_b200 as v

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this library has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist = None

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    w = args.workload
    size = SIZES[w]
    os.environ.setdefault("VKVG_B200_DEVICE", str(local))
    dev = v.Device(4, analytic=args.coverage == "analytic")
    striped = w == "c5a"
    if striped:   # strong scaling: every rank replays the same scene onto its own tile-row stripe
        from vkvg_b200 import sharding
        y0, sh = sharding.stripe_rows(size, world)[rank]
        surf = v.Surface(dev, size, sh, full_height=size, origin_y=y0)
        emit, units, info = build_scene(w, 1, args.rule)
    elif w == "c5b":   # strong scaling: the 1024 canvases are split across the ranks, C5B_BATCH canvases per flush
        surf = v.Surface(dev, size, size, batch=C5B_BATCH)
        emit, units, info = build_scene(w, 1 + rank, args.rule)
        assert (C5B_CANVASES // C5B_BATCH) % world == 0, "world size must divide %d batches" % (C5B_CANVASES // C5B_BATCH)
        reps = C5B_CANVASES // C5B_BATCH // world
        units = float(C5B_CANVASES)
    else:
        surf = v.Surface(dev, size, size)
        emit, units, info = build_scene(w, 1 + rank, args.rule)
    strong = striped or w == "c5b"
    if w != "c5b":
        reps = 1
    ctx = v.Context(surf)
    cs = v.CommandStream()
    direct = None
    if w == "blit":   # surface sources are handles, not numbers: this workload drives the C API call by call instead of a command stream
        img = layer_image()
        hlayer = v.lib().vkvg_surface_create_from_bitmap(dev.h, img.ctypes.data, 2048, 2048)
        layer = v.Surface.__new__(v.Surface)
        layer.dev, layer.width, layer.height, layer.full_height, layer.origin_y, layer.batch, layer.h = dev, 2048, 2048, 2048, 0, None, hlayer
        direct = LayerSource(ctx, layer)
    else:
        emit(cs)
    ops_np, args_np = cs.arrays()
    # host buffers of the end-to-end path live in pinned memory
    ops_t = torch.from_numpy(ops_np.copy()).pin_memory()
    args_t = torch.from_numpy(args_np.copy()).pin_memory()
    out_t = torch.empty((surf.height, size, 4), dtype=torch.uint8).pin_memory()
    L = v.lib()
    gather_ms = [0.0]

    def gather():   # striped surfaces only: reassemble on every rank with one NCCL all-gather of contiguous rows
        if not striped or dist is None:
            return None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        local_t = torch.empty((surf.height, size, 4), dtype=torch.uint8, device="cuda")
        surf.copy_to_device(local_t.data_ptr())
        e0.record()
        full = sharding.gather_stripes(local_t, size)
        e1.record()
        torch.cuda.synchronize()
        gather_ms[0] += e0.elapsed_time(e1)
        return full

    parts = [0.0] * 5

    def e2e_step():
        t0 = time.perf_counter()
        L.vkvg_clear(ctx.h)
        if direct is not None:
            emit(direct)
            ctx.identity_matrix()
            ctx.set_opacity(1.0)
            st = 0
        else:
            st = L.vkvg_b200_replay(ctx.h, ops_t.data_ptr(), ops_t.numel(), args_t.data_ptr(), args_t.numel())
        assert st == 0, st
        t1 = time.perf_counter()
        L.vkvg_flush(ctx.h)   # queues the upload and the whole pipeline (a CUDA graph replay once the frame structure repeats) and returns
        t2 = time.perf_counter()
        gather()
        assert L.vkvg_b200_surface_read_premultiplied(surf.h, out_t.data_ptr()) == 0
        t3 = time.perf_counter()
        parts[0] += t1 - t0
        parts[1] += t2 - t1
        parts[2] += t3 - t2

    for _ in range(max(args.warmup, 3)):
        e2e_step()
    e2e_one = e2e_step

    def e2e_step():   # noqa: F811  (a step covers every batch this rank owns)
        for _ in range(reps):
            e2e_one()
    checksum = int(out_t.numpy().view(np.uint32).sum(dtype=np.uint64))
    dev.set_profiling(True)
    dev.set_stage_timing(True)
    dev.time_resident(surf, 2, True, True)   # warm the resident path (buffers sized, L2 scratch allocated)
    # ---- per-stage breakdown: plain launches with CUDA events between the stages (not the headline timing) ----
    st_stages = dev.time_resident(surf, args.steps * reps, True, True)
    dev.set_stage_timing(False)
    use_graph = not args.no_graph
    dev.set_graphs(use_graph)
    dev.time_resident(surf, 4, True, True)   # the third flush of a given structure captures the CUDA graph, later ones replay it

    sampler = ClockSampler(local)
    sampler.start()
    # ---- device-timed, inputs resident in HBM: the whole flush as one CUDA graph replay per step ----
    barrier()
    l0 = L.vkvg_b200_launch_count()
    g0 = dev.graph_replays()
    st = dev.time_resident(surf, args.steps * reps, True, True)
    launches = L.vkvg_b200_launch_count() - l0
    graph_replays = dev.graph_replays() - g0
    gather_ms[0] = 0.0
    for _ in range(args.steps if striped else 0):
        gather()
    barrier()
    ms_step = max_over_ranks((st["ms_total"] + gather_ms[0]) / args.steps)
    gather_step_ms = gather_ms[0] / args.steps
    # ---- end to end through the C ABI with host buffers ----
    dev.set_profiling(False)   # flushes return as soon as the work is queued; the read-back waits for it
    for _ in range(3):
        e2e_step()
    barrier()
    parts[:] = [0.0] * 5
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
    clocks = sampler.result()
    assert int(out_t.numpy().view(np.uint32).sum(dtype=np.uint64)) == checksum, "non-deterministic output"

    # ---- roofline of the dominant kernel (fine pass: winding + paint + OVER + resolve, one launch per step) ----
    peak, peak_src = measured_peak_hbm()
    n_edges, n_draws = st["n_edges"], info["n_paths"]
    alg_bytes = 16 * n_edges + 32 * n_draws + 4 * size * surf.height
    if w == "blit":   # + every source texel a layer covers, read once per layer (4 B x 2048^2 x 8 layers)
        alg_bytes += 4 * 2048 * 2048 * 8
    # the fine kernel's duration: CUDA events recorded around it inside the timed region (external event nodes of the replayed
    # graph); if the driver did not time those, the per-stage pass above (same kernel, plain launch) supplies it
    fine_src = "events around the kernel inside the timed graph replays"
    fine_ms = st["ms_fine"] / (args.steps * reps)   # per launch
    if not fine_ms > 0:
        fine_ms = st_stages["ms_fine"] / (args.steps * reps)
        fine_src = "events around the kernel in %d plain-launch steps run before the timed region" % args.steps
    achieved = alg_bytes / (fine_ms * 1e-3) / 1e9
    stage = {k: val / (args.steps * reps) for k, val in st_stages["ms_stage"].items()}   # per flush
    # which fine kernel ran: batches without clip state go to the warp-per-tile kernel from 16384 tiles up (raster.cu: FW_MIN_TILES)
    n_tiles = ((size + 15) // 16) * ((surf.height + 15) // 16)
    mode = L.vkvg_b200_get_fine_kernel()
    if args.coverage != "msaa":
        fine_name = "fine_analytic_k"
    elif mode == 2 or (mode == 0 and n_tiles >= 16384):
        fine_name = "fine_warp_k<4>"
    else:
        fine_name = "fine_k<4>"
    traffic = None   # DRAM bytes of the dominant kernel per launch, from the committed ncu --set full capture of this workload
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        for key in (w, w + "_block"):
            if key in tj and tj[key]["kernel"] == fine_name and world == 1:
                traffic = int(tj[key]["dram_bytes_per_launch"])
    except Exception:
        pass
    name, unit = UNITS[w]
    line = {
        "metric": name, "value": (1 if strong else world) * units / (ms_step * 1e-3), "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32+i64", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAMES[w] if args.coverage == "msaa" else WORKLOAD_NAMES[w].replace("4 samples", "analytic coverage"),
                   "rule": args.rule, "samples": 4 if args.coverage == "msaa" else 0, "coverage": args.coverage, "sharding": ("tile-row stripes of one surface, %.3f ms all-gather per step" % gather_step_ms) if striped else ("%d canvases per rank in %d flushes of %d" % (C5B_CANVASES // world, reps, C5B_BATCH) if w == "c5b" else "one independent canvas per rank"),
                   "l2": "256 MiB scratch overwritten between timed steps", "launch": ("one CUDA graph replay per flush (%d of %d flushes)" % (graph_replays, args.steps * reps)) if use_graph else "plain kernel launches",
                   **info, "n_edges": int(n_edges),
                   "n_tile_edges": int(st["n_tile_edges"]), "n_points": int(st["n_points"]), "n_path_tiles": int(st["n_nonempty"])},
        "e2e": {"value": (1 if strong else world) * units / e2e_s, "unit": unit, "h2d_bytes_per_step": int(ops_t.numel() + 4 * args_t.numel()) * reps,
                "d2h_bytes_per_step": int(out_t.numel()) * reps, "ms_per_step": e2e_s * 1e3,
                "host_record_ms": parts[0] / args.steps * 1e3, "upload_render_ms": parts[1] / args.steps * 1e3,
                "readback_ms": parts[2] / args.steps * 1e3, "h2d_bytes_wire": int(st["h2d_bytes"])},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": fine_name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "algorithmic_bytes": int(alg_bytes), "kernel_ms": fine_ms, "kernel_ms_source": fine_src, "peak_source": peak_src,
                     "whole_step_frac": alg_bytes * reps / (ms_step * 1e-3) / 1e9 / peak},
        "stage_ms": stage, "stage_ms_note": "per flush, plain launches with events between stages: %.3f ms per flush" % (st_stages["ms_total"] / (args.steps * reps)),
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(w, args.rule)
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def cpu_baseline(workload, rule):
    """bounded sample of the same workload on ONE host core: reference tessellation object code (oracle/_ref) when it was
    built, else the oracle port; rasterised by the oracle's scalar Vulkan restatement."""
    import oracle
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        oracle.build(ref=False)
    n = CPU_SAMPLE[workload]
    dt, info, kind = _ref_worker(("c1" if workload == "c5b" else workload, 1, rule, n, 0))
    frac = (1.0 / C5B_CANVASES) if workload == "c5b" else (1.0 if n is None else n / FULL[workload])
    name, unit = UNITS[workload]
    return {"value": build_units(workload) * frac / dt, "unit": unit, "cores": 1, "kind": kind, "seconds": dt,
            "sample": ("the whole scene" if n is None else "first %d of %d units of the scene" % (n, FULL[workload])) +
                      ", full-size surface, 1 thread; throughput scaled by the fraction processed"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(SIZES))
    ap.add_argument("--rule", default="nz", choices=["nz", "eo"])
    ap.add_argument("--coverage", default="msaa", choices=["msaa", "analytic"],
                    help="msaa: 4-sample mode, bit-exact with the reference's rasterisation (default); analytic: exact-area coverage")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time plain kernel launches instead of CUDA graph replays")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

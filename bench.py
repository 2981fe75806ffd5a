#!/usr/bin/env python
"""bench.py — the driver's benchmark contract for the vkvg path-rendering hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c3|c4|c5a|c5b|blit] [--rule nz|eo] [--coverage msaa|analytic]
                    [--no-graph] [--impl reference] [--only]

One "step" = one pass of the hot path (flatten -> stroke / fill edges -> tile binning -> winding + paint + OVER) over one synthetic
scene.  The headline line is BASELINE.json configs[1] ("c2": 100k random self-intersecting polygons, non-zero, 4096x4096, 4 samples); at
N > 1 every rank renders its own independent canvas of that configuration (weak scaling, no collective on the data path: SURVEY.md 8e).
The same line carries, unless --only is given:
  configs  the other single-GPU configurations of BASELINE.json measured the same way on rank 0's GPU: c1 (tiger frames/s, with
           vkvg_surface_write_to_png timed separately), c2_eo, c3 (Msegments/s), c4
  sharded  the two configurations that shard, STRONG scaling over the N ranks of this run: c5a (tile-row stripes of one
           16384^2 surface, each stripe delivered into rank 0's picture over NVLink while it renders) and c5b (1024 independent tiger canvases split across the ranks)
  parity   the GPU's C2 frame against the cpu_baseline leg's frame (reference tessellation + oracle raster), every pixel

  value  = whole-job throughput with the recorded scene already resident in HBM (device-timed: CUDA events on the library's stream around
           clear + whole flush, max over ranks, L2 flushed between steps outside the timed events); each step replays the CUDA graph the
           library captures for a repeating frame (--no-graph: plain launches); a pass with plain launches supplies stage_ms
  e2e    = the same metric through the public C ABI with HOST buffers: packed command arrays in pinned host memory -> one submit call ->
           flush -> the finished surface in pinned host memory, all inside the timed region (wall clock, max over ranks)
  --impl reference: the reference's own CPU implementation of the path on the host cores: its unmodified tessellation object code
           (oracle/_ref) + the oracle's scalar restatement of the Vulkan rasteriser it delegates to ("restated CPU baseline - not
           lavapipe", BASELINE.md 5.2: no Vulkan ICD exists in this image or on the GPU box).  Every step renders the WHOLE scene: the
           surface is cut into one band of rows per host process, each process draws the paths that can reach its band.
"""
import argparse
import json
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
    "blit": "layer compositing: 8 translucent 2048x2048 surface sources (bilinear, rotated) painted over a 4096x4096 surface (SURVEY 8f rank 2)",
    "c5a": "C5a 16384x16384 surface, 5M-segment mix (C2-style polygons + C3-style dashed polylines), sharded by tile-row stripes, stripes delivered into rank 0's picture",
}
REF_LABEL = "restated CPU baseline - not lavapipe (reference tessellation object code + scalar restatement of the Vulkan rasteriser)"
BIG = 1e30


# ---------------------------------------------------------------------------------------------------------------
# scenes: (setup(g), [(ymin, ymax, draw(g))], units, info).  `g` is any object with the drawing vocabulary (the CUDA library's context or
# command stream, the oracle, the reference).  ymin / ymax bound the rows an item can touch (the reference arm draws an item only in
# the bands it reaches); items are in submission order.
# ---------------------------------------------------------------------------------------------------------------
def poly(g, pts):
    if hasattr(g, "polyline"):
        g.polyline(pts)
    else:
        g.move_to(float(pts[0, 0]), float(pts[0, 1]))
        for p in pts[1:]:
            g.line_to(float(p[0]), float(p[1]))


def _fill_item(p, c):
    def draw(g):
        g.set_source_rgba(*c)
        poly(g, p)
        g.close_path()
        g.fill()
    return (float(p[:, 1].min()), float(p[:, 1].max()), draw)


def _stroke_setup(g):
    g.set_line_width(3.0)
    g.set_line_join(1)
    g.set_line_cap(1)
    g.set_dash([10.0, 6.0], 0.0)


def scene_items(workload, seed, rule, n_limit=None, first=0):
    from tests import scenes
    size = SIZES[workload]
    if workload == "c2":
        polys, cols = scenes.polygons_c2(100000, size, seed)
        sl = slice(first, first + n_limit) if n_limit else slice(None)
        polys, cols = polys[sl], cols[sl]
        items = [_fill_item(p, [float(x) for x in c]) for p, c in zip(polys, cols)]
        return (lambda g: g.set_fill_rule(0 if rule == "eo" else 1)), items, size * size / 1e6, dict(n_paths=len(polys), n_segments=int(sum(len(p) for p in polys)))
    if workload == "c3":
        n = 1_000_001 if not n_limit else n_limit + 1
        pts = scenes.polyline_c3(n, size, seed)

        def setup(g):
            g.set_source_rgba(0.1, 0.2, 0.8, 1.0)
            _stroke_setup(g)

        def draw(g):
            poly(g, pts)
            g.stroke()
        return setup, [(-BIG, BIG, draw)], (n - 1) / 1e6, dict(n_paths=1, n_segments=n - 1)
    if workload == "c4":
        paths = scenes.curves_c4(50000, size, seed)
        sl = slice(first, first + n_limit) if n_limit else slice(None)
        paths = paths[sl]

        def item(pts, kind, cx, cy, stops):
            def draw(g):
                if kind == 0:
                    g.set_source_linear(cx - 64, cy - 64, cx + 64, cy + 64, stops)
                else:
                    g.set_source_radial(cx, cy, 4.0, cx + 8, cy - 8, 96.0, stops)
                g.move_to(float(pts[-1, 2, 0]), float(pts[-1, 2, 1]))
                for s in pts:
                    g.curve_to(*[float(x) for x in s.ravel()])
                g.close_path()
                g.fill()
            return (float(pts[..., 1].min()), float(pts[..., 1].max()), draw)   # (a cubic stays inside the hull of its control points)
        items = [item(pts, kind, cx, cy, stops) for pts, kind, (cx, cy), stops in paths]
        return (lambda g: g.set_fill_rule(1)), items, size * size / 1e6, dict(n_paths=len(paths), n_segments=int(sum(len(p[0]) for p in paths)))
    if workload == "c5a":
        n_poly = 238000 if not n_limit else n_limit
        polys, cols = scenes.polygons_c2(n_poly, size, seed)
        # polygons_c2 draws radii for a 4096 canvas; keep them (small shapes on a huge surface), plus 25 polylines of 100k segments
        lines = [scenes.polyline_c3(100001, size, seed * 100 + i) for i in range(25 if not n_limit else 1)]
        items = [_fill_item(p, [float(x) for x in c]) for p, c in zip(polys, cols)]

        def line_item(i, pts):
            def draw(g):
                if i == 0:
                    _stroke_setup(g)
                g.set_source_rgba(0.1 + 0.03 * i, 0.2, 0.8 - 0.02 * i, 1.0)
                poly(g, pts)
                g.stroke()
            return (-BIG, BIG, draw)   # the dash phase depends on the whole prefix: a polyline is always stroked completely
        items += [line_item(i, pts) for i, pts in enumerate(lines)]
        nseg = int(sum(len(p) for p in polys)) + sum(len(ln) - 1 for ln in lines)
        return (lambda g: g.set_fill_rule(1)), items, size * size / 1e6, dict(n_paths=len(polys) + len(lines), n_segments=nseg)
    if workload == "blit":
        n_layers = n_limit or 8

        def draw(g):
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
        return (lambda g: None), [(-BIG, BIG, draw)], size * size / 1e6, dict(n_paths=n_layers + 1, n_segments=4 * n_layers)
    if workload == "c5b":   # one flush worth of canvases; the step replays it for every batch this rank owns
        w, h, shapes = scenes.load_nsvg(os.path.join(ROOT, "tests", "golden", "tiger.nsvg.bin"))
        r = scenes.SplitMix64(900 + seed)
        jit = [(r.uniform(-2, 2), r.uniform(-2, 2)) for _ in range(C5B_BATCH)]

        def draw(g):
            for i, (jx, jy) in enumerate(jit):
                if hasattr(g, "set_canvas"):
                    g.set_canvas(i)
                g.identity_matrix()
                g.translate(jx, jy)
                scenes.render_nsvg(g, shapes)
        nseg = int(sum((len(p) - 1) // 3 for s in shapes for p, _ in s["paths"])) * C5B_BATCH
        return (lambda g: None), [(-BIG, BIG, draw)], float(C5B_BATCH), dict(n_paths=len(shapes) * C5B_BATCH, n_segments=nseg)
    if workload == "c1":
        w, h, shapes = scenes.load_nsvg(os.path.join(ROOT, "tests", "golden", "tiger.nsvg.bin"))
        if n_limit:
            shapes = shapes[first:first + n_limit]
        return (lambda g: None), [(-BIG, BIG, lambda g: scenes.render_nsvg(g, shapes))], 1.0, dict(
            n_paths=len(shapes), n_segments=int(sum((len(p) - 1) // 3 for s in shapes for p, _ in s["paths"])))
    raise SystemExit("unknown workload " + workload)


def build_scene(workload, seed, rule, n_limit=None, first=0):
    """returns (emit(g), units, info): emit replays the whole scene on g; units = the metric's unit count of one step."""
    setup, items, units, info = scene_items(workload, seed, rule, n_limit, first)

    def emit(g):
        setup(g)
        for _, _, draw in items:
            draw(g)
    return emit, units, info


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


def build_units(workload):
    return {"blit": 4096 * 4096 / 1e6, "c5b": float(C5B_CANVASES), "c1": 1.0, "c2": 4096 * 4096 / 1e6, "c3": 1.0, "c4": 8192 * 8192 / 1e6, "c5a": 16384 * 16384 / 1e6}[workload]


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
        except Exception:
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md); MEASURED_PEAKS.json absent"


# ---------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: reference tessellation object code + oracle raster, on the host cores
# ---------------------------------------------------------------------------------------------------------------
_SCENES = {}   # (workload, seed, rule) -> scene_items(...): built in the parent before the pool forks, inherited by the workers


def _render_cpu(workload, setup, items, window=None):
    """draw `items` through the reference's object code (oracle/_ref; the oracle's port where a workload needs what the recording shim does
    not carry) and rasterise with the oracle; returns (seconds of rendering, pixels, kind)."""
    import oracle
    from oracle import Oracle, Ref
    size = SIZES[workload]
    kind = "reference" if oracle.ref_available() else "port"
    if workload == "blit":   # textures do not travel through the recorded draw list of the reference build: oracle port only
        kind = "port"
    t0 = time.perf_counter()
    o = Oracle(size, size, 4, window=window)
    g = Ref(size, size, 4) if kind == "reference" else o
    if workload == "blit":
        g = LayerSource(g, layer_image())
    setup(g)
    for _, _, draw in items:
        draw(g)
    if kind == "reference":
        g.render_with(o)   # the reference's recorded draw list, rasterised by the Vulkan restatement
        g.close()
    px = o.pixels()
    dt = time.perf_counter() - t0
    o.close()
    return dt, px, kind


def _band_worker(args):
    """one band of rows of one scene: the items that can reach it, in order, onto a window of the logical surface"""
    key, band, n_bands = args
    setup, items, _, _ = _SCENES[key]
    size = SIZES[key[0]]
    rows = (size + 15) // 16
    y0, y1 = min(size, 16 * (rows * band // n_bands)), min(size, 16 * (rows * (band + 1) // n_bands))
    if y1 <= y0:
        return 0.0, "reference", 0
    mine = [it for it in items if it[1] >= y0 - 2 and it[0] <= y1 + 2]
    dt, _, kind = _render_cpu(key[0], setup, mine, window=(0, y0, size, y1 - y0))
    return dt, kind, len(mine)


def _ref_worker(args):
    """a bounded slice of a scene on the whole surface, one process (tests, --workload X cpu_baseline legs)"""
    workload, seed, rule, n_limit, first = args
    setup, items, _, info = scene_items(workload, seed, rule, n_limit=n_limit, first=first)
    dt, _, kind = _render_cpu(workload, setup, items)
    return dt, info, kind


# units of work of the single-thread cpu_baseline leg (about 10-30 s of CPU work): None = the whole scene
CPU_SAMPLE = {"blit": 1, "c5b": None, "c1": None, "c2": None, "c3": None, "c4": 2000, "c5a": 4000}
SAMPLE = CPU_SAMPLE   # (kept for tests/test_bench_cpu.py)
FULL = {"blit": 8, "c5b": 239 * C5B_CANVASES, "c1": 239, "c2": 100000, "c3": 1000000, "c4": 50000, "c5a": 476000}


def run_reference(args):
    """--impl reference.  Every step renders `n_scenes` WHOLE scenes of the workload (one at N = 1; at N > 1 as many canvases as the GPU arm
    renders per step, seeds 1..N, so that the driver's ratio compares equal work): each scene is cut into one band of tile rows per
    host process; a band process draws every path that can reach its rows (paths that span several bands are tessellated in each, as
    a tiled CPU rasteriser would) onto a window of the logical surface.  Time = wall clock around the whole step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    import oracle
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        oracle.build(ref=False)
    cores = os.cpu_count() or 1
    w = args.workload
    strong = w in ("c5a", "c5b")
    n_scenes = 1 if strong else max(1, args.gpus)
    base = "c1" if w == "c5b" else w
    keys = [(base, 1 + i, args.rule) for i in range(n_scenes)]
    for k in keys:
        _SCENES[k] = scene_items(k[0], k[1], k[2])
    if w == "c5b":     # independent canvases: a step = one whole tiger canvas per host process (no bands), value in canvases/s
        jobs = [(keys[0], 0, 1) for _ in range(cores)]
        units_step, what = float(cores), "%d whole tiger canvases per step, one per host process" % cores
    else:
        jobs = [(k, b, cores) for k in keys for b in range(cores)]
        units_step = build_units(w) * n_scenes
        what = "%d whole scene%s per step, %d bands of tile rows, one host process each" % (n_scenes, "s" if n_scenes > 1 else "", cores)
    kind = "port"
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(min(args.warmup, 1)):
            pool.map(_band_worker, jobs, chunksize=1)
        ts = []
        for _ in range(args.steps):
            t0 = time.perf_counter()
            res = pool.map(_band_worker, jobs, chunksize=1)
            ts.append(time.perf_counter() - t0)
            kind = res[0][1]
    t = float(np.mean(ts))
    value = units_step / t
    name, unit = UNITS[w]
    line = {"impl": "reference", "metric": name, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1),
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32+i64", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAMES[w], "rule": args.rule, "samples": 4},
            "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": kind, "label": REF_LABEL, "sample": what,
                             "busiest_band_s": float(max(r[0] for r in res)), "mean_band_s": float(np.mean([r[0] for r in res]))},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def cpu_baseline(workload, rule):
    """the same workload on ONE host core: reference tessellation object code (oracle/_ref) when it was built, else the oracle port,
    rasterised by the oracle's scalar Vulkan restatement on the whole surface.  Returns (record, pixels or None)."""
    import oracle
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        oracle.build(ref=False)
    n = CPU_SAMPLE[workload]
    base = "c1" if workload == "c5b" else workload
    setup, items, _, info = scene_items(base, 1, rule, n_limit=n)
    dt, px, kind = _render_cpu(base, setup, items)
    frac = (1.0 / C5B_CANVASES) if workload == "c5b" else (1.0 if n is None else n / FULL[workload])
    name, unit = UNITS[workload]
    rec = {"value": build_units(workload) * frac / dt, "unit": unit, "cores": 1, "kind": kind, "label": REF_LABEL, "seconds": dt,
           "sample": ("the whole scene" if n is None else "first %d of %d units of the scene" % (n, FULL[workload])) +
                     ", full-size surface, 1 thread" + ("" if n is None else "; throughput scaled by the fraction processed")}
    return rec, (px if n is None else None)


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
class Env:
    """process-wide state of one bench run: torch.distributed plumbing and the CUDA library's device"""

    def __init__(self, args):
        import torch
        import vkvg_b200 as v
        self.torch, self.v, self.args = torch, v, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; this library has no CPU path (use --impl reference for the CPU baseline)")
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        os.environ.setdefault("VKVG_B200_DEVICE", str(self.local))
        self.L = v.lib()

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def measure(env, w, rule, steps, strong_world=None):
    """one configuration on this rank's GPU -> its record (every number a max over the ranks that took part).
    strong_world: None = every rank renders its own whole canvas (weak); else the scene is split over that many ranks (c5a, c5b)."""
    torch, v, L, args = env.torch, env.v, env.L, env.args
    rank, world = env.rank, env.world
    # a device object of its own per configuration: the library sizes its launches from the largest batch a device has seen
    # (grow-only capacities), so a small scene measured after a large one on the same device object would pay for the large one's grids
    dev = v.Device(4, analytic=args.coverage == "analytic")
    size = SIZES[w]
    striped = w == "c5a"
    strong = striped or w == "c5b"
    reps = 1
    if striped:   # strong scaling: every rank replays the same scene onto its own tile-row stripe
        from vkvg_b200 import sharding
        surf, y0, sh = sharding.stripe_surface(dev, size, size, rank, world)
        emit, units, info = build_scene(w, 1, rule)
    elif w == "c5b":   # strong scaling: the 1024 canvases are split across the ranks, C5B_BATCH canvases per flush
        surf = v.Surface(dev, size, size, batch=C5B_BATCH)
        emit, units, info = build_scene(w, 1 + rank, rule)
        assert (C5B_CANVASES // C5B_BATCH) % world == 0, "world size must divide %d batches" % (C5B_CANVASES // C5B_BATCH)
        reps = C5B_CANVASES // C5B_BATCH // world
        units = float(C5B_CANVASES)
    else:
        surf = v.Surface(dev, size, size)
        emit, units, info = build_scene(w, 1 + (rank if strong_world is None else 0), rule)
    ctx = v.Context(surf)
    cs = v.CommandStream()
    direct = None
    if w == "blit":   # surface sources are handles, not numbers: this workload drives the C API call by call instead of a command stream
        img = layer_image()
        hlayer = L.vkvg_surface_create_from_bitmap(dev.h, img.ctypes.data, 2048, 2048)
        layer = v.Surface.__new__(v.Surface)
        layer.dev, layer.width, layer.height, layer.full_height, layer.origin_y, layer.batch, layer.h = dev, 2048, 2048, 2048, 0, None, hlayer
        direct = LayerSource(ctx, layer)
    else:
        emit(cs)
    ops_np, args_np = cs.arrays2() if direct is None else (np.zeros(0, np.uint32), np.zeros(0, np.float32))
    # host buffers of the end-to-end path live in pinned memory: the command stream (cmds = op | n_args << 8, args) and the image
    ops_t = torch.from_numpy(ops_np.view(np.int32).copy()).pin_memory()
    args_t = torch.from_numpy(args_np.copy()).pin_memory()
    out_t = torch.empty((surf.height, size, 4), dtype=torch.uint8).pin_memory()
    on_dev0 = v.submit_counts()
    gather_ms = [0.0]
    full_t = [None]
    # stripes on more than one GPU: every rank's finished bands go straight into the root's picture over NVLink (CUDA IPC pointer as the
    # read-back target of the stripe surface, copy engines, overlapped with the bands still rendering) - no collective behind the frame
    peer = sharding.deliver_to_root(surf, y0, sh, size, size) if striped and env.dist is not None and world > 1 else None

    def gather():   # striped surfaces only: the finished stripes are gathered into one image on rank 0 (NCCL over NVLink)
        if not striped or env.dist is None:
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        full_t[0] = sharding.gather_surface_to_root(surf, size, out=full_t[0])
        e1.record()
        torch.cuda.synchronize()
        gather_ms[0] += e0.elapsed_time(e1)

    parts = [0.0] * 3

    def e2e_one():
        t0 = time.perf_counter()
        L.vkvg_clear(ctx.h)
        if direct is not None:
            emit(direct)
            ctx.identity_matrix()
            ctx.set_opacity(1.0)
            st = 0
            t1 = time.perf_counter()
            L.vkvg_flush(ctx.h)
        else:
            # one call: the stream is uploaded and decoded into the batch by kernels (decode.cu), the pipeline is queued behind it - a CUDA
            # graph replay once the frame structure repeats - and the call returns; finished bands of the image are copied to out_t on a
            # second stream while later bands render (vkvg_b200_surface_set_readback)
            st = L.vkvg_b200_submit(ctx.h, ops_t.data_ptr(), ops_t.numel(), args_t.data_ptr(), args_t.numel())
            t1 = time.perf_counter()
        assert st == 0, st
        t2 = time.perf_counter()
        if peer is None:
            gather()
        assert L.vkvg_b200_surface_read_premultiplied(surf.h, out_t.data_ptr()) == 0
        t3 = time.perf_counter()
        parts[0] += t1 - t0
        parts[1] += t2 - t1
        parts[2] += t3 - t2

    def e2e_step():   # a step covers every batch this rank owns
        for _ in range(reps):
            e2e_one()

    if not striped:
        L.vkvg_b200_surface_set_readback(surf.h, out_t.data_ptr())
    for _ in range(3):
        e2e_one()
    frame = out_t.numpy().copy()
    checksum = int(frame.view(np.uint32).sum(dtype=np.uint64))
    if peer is None:
        L.vkvg_b200_surface_set_readback(surf.h, None)   # (the device-timed runs below render only)
    else:   # (... and deliver: the root's picture must hold every rank's rows)
        peer.barrier()
        sums = torch.tensor([checksum if sh > 0 else 0, 0], dtype=torch.int64, device="cuda")
        if rank == 0:
            sums[1] = peer.full.as_tensor().view(torch.int32).to(torch.int64).bitwise_and(0xffffffff).sum()
        env.dist.all_reduce(sums)
        assert int(sums[0]) == int(sums[1]), "the root's picture is not the sum of the stripes: %d != %d" % (int(sums[0]), int(sums[1]))
    dev.set_profiling(True)
    dev.set_stage_timing(True)
    dev.time_resident(surf, 2, True, True)   # warm the resident path (buffers sized, L2 scratch allocated)
    # ---- per-stage breakdown: plain launches with CUDA events between the stages (not the headline timing) ----
    st_stages = dev.time_resident(surf, steps * reps, True, True)
    dev.set_stage_timing(False)
    use_graph = not args.no_graph
    dev.set_graphs(use_graph)
    dev.time_resident(surf, 4, True, True)   # the third flush of a given structure captures the CUDA graph, later ones replay it

    sampler = ClockSampler(env.local)
    sampler.start()
    # ---- device-timed, inputs resident in HBM: clear + the whole flush as one CUDA graph replay per step ----
    env.barrier()
    l0 = L.vkvg_b200_launch_count()
    g0 = dev.graph_replays()
    st = dev.time_resident(surf, steps * reps, True, True)
    launches = L.vkvg_b200_launch_count() - l0
    graph_replays = dev.graph_replays() - g0
    env.barrier()   # (the ranks leave the render loop at different times: the first gather must not be charged for the wait)
    gather_ms[0] = 0.0
    notify_ms = 0.0
    if peer is not None:
        # the frame time above already holds the delivery (the stream that renders joins the copy stream before the end event); what a
        # consumer on the root still needs is the ranks' word that their rows are there: one barrier per frame, timed on its own
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            env.dist.barrier()
        torch.cuda.synchronize()
        notify_ms = (time.perf_counter() - t0) / 20 * 1e3
        peer.surf.set_readback(None)      # the NCCL gather of the same stripes behind a render-only frame, for comparison
        dev.time_resident(surf, 4, True, True)   # (another read-back target is another graph: captured here)
        env.barrier()
        render_only = dev.time_resident(surf, steps, True, True)
        gather()   # (the first one sets up NCCL's point-to-point channels)
        env.barrier()
        gather_ms[0] = 0.0
        for _ in range(steps):
            gather()
        peer.surf.set_readback(peer.target or None)
        dev.time_resident(surf, 4, True, True)
        env.barrier()
        ms_step = env.max_over_ranks(st["ms_total"] / steps) + env.max_over_ranks(notify_ms)
        gather_step_ms = env.max_over_ranks(gather_ms[0] / steps)
        render_only_ms = env.max_over_ranks(render_only["ms_total"] / steps)
    else:
        for _ in range(steps if striped else 0):
            gather()
        env.barrier()
        ms_step = env.max_over_ranks((st["ms_total"] + gather_ms[0]) / steps)
        gather_step_ms = env.max_over_ranks(gather_ms[0] / steps)
    # ---- end to end through the C ABI with host buffers ----
    dev.set_profiling(False)   # flushes return as soon as the work is queued; the read-back waits for it
    if not striped:
        L.vkvg_b200_surface_set_readback(surf.h, out_t.data_ptr())
    for _ in range(2):
        e2e_step()
    env.barrier()
    parts[:] = [0.0] * 3
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    env.barrier()
    e2e_s = env.max_over_ranks((time.perf_counter() - t0) / steps)
    clocks = sampler.result()
    assert int(out_t.numpy().view(np.uint32).sum(dtype=np.uint64)) == checksum, "non-deterministic output"

    # ---- roofline of the dominant kernel (fine pass: winding + paint + OVER + resolve, one launch per flush) ----
    peak, peak_src = measured_peak_hbm()
    n_edges, n_draws = st["n_edges"], info["n_paths"]
    alg_bytes = 16 * n_edges + 32 * n_draws + 4 * size * surf.height
    if w == "blit":   # + every source texel a layer covers, read once per layer (4 B x 2048^2 x 8 layers)
        alg_bytes += 4 * 2048 * 2048 * 8
    fine_src = "events around the kernel inside the timed graph replays"
    fine_ms = st["ms_fine"] / (steps * reps)   # per launch
    if not fine_ms > 0:
        fine_ms = st_stages["ms_fine"] / (steps * reps)
        fine_src = "events around the kernel in %d plain-launch steps run before the timed region" % steps
    achieved = alg_bytes / (fine_ms * 1e-3) / 1e9
    stage = {k: val / (steps * reps) for k, val in st_stages["ms_stage"].items()}   # per flush
    n_tiles = ((size + 15) // 16) * ((surf.height + 15) // 16)
    mode = L.vkvg_b200_get_fine_kernel()
    if args.coverage != "msaa":
        fine_name = "fine_analytic_k"
    elif mode == 2 or (mode == 0 and n_tiles >= 16384):
        fine_name = "fine_warp_k<4>"
    else:
        fine_name = "fine_k<4>"
    traffic = issue = None   # DRAM bytes / issue-slot utilisation of the dominant kernel, from the committed ncu --set full capture of this workload
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        for key in (w, w + "_block"):
            if key in tj and tj[key]["kernel"] == fine_name and world == 1:
                traffic = int(tj[key]["dram_bytes_per_launch"])
                issue = tj[key].get("issue_slots_busy")
    except Exception:
        pass
    name, unit = UNITS[w]
    scale_units = 1 if strong else (world if strong_world is None else 1)
    rec = {
        "metric": name, "value": scale_units * units / (ms_step * 1e-3), "unit": unit, "ms_per_step": ms_step, "scaling": "strong" if strong else "weak",
        "config": {"workload": WORKLOAD_NAMES[w] if args.coverage == "msaa" else WORKLOAD_NAMES[w].replace("4 samples", "analytic coverage"),
                   "rule": rule, "samples": 4 if args.coverage == "msaa" else 0, "coverage": args.coverage,
                   "sharding": (("tile-row stripes of one surface over %d ranks, finished bands delivered into rank 0's picture over NVLink while later bands render "
                                 "(+ %.3f ms barrier per frame); render-only frame %.3f ms, NCCL gather behind it %.3f ms" % (world, notify_ms, render_only_ms, gather_step_ms)) if peer is not None
                                else "tile-row stripes of one surface over %d ranks, %.3f ms gather to rank 0 per step" % (world, gather_step_ms)) if striped else (
                       "%d canvases per rank in %d flushes of %d" % (C5B_CANVASES // world, reps, C5B_BATCH) if w == "c5b" else "one independent canvas per rank"),
                   "l2": "256 MiB scratch overwritten between timed steps",
                   "launch": ("one CUDA graph replay per flush (%d of %d flushes)" % (graph_replays, steps * reps)) if use_graph else "plain kernel launches",
                   **info, "n_edges": int(n_edges), "n_tile_edges": int(st["n_tile_edges"]), "n_points": int(st["n_points"]), "n_path_tiles": int(st["n_nonempty"])},
        "e2e": {"value": scale_units * units / e2e_s, "unit": unit, "h2d_bytes_per_step": int(4 * ops_t.numel() + 4 * args_t.numel()) * reps,
                "decoded_on": "device" if v.submit_counts()[0] > on_dev0[0] and v.submit_counts()[1] == on_dev0[1] else ("host" if direct is None else "host (API calls)"),
                "d2h_bytes_per_step": int(out_t.numel()) * reps, "ms_per_step": e2e_s * 1e3,
                "host_record_ms": parts[0] / steps * 1e3, "upload_render_ms": parts[1] / steps * 1e3,   # submit call (upload + decode + queue) | explicit flush (call-by-call workloads only)
                "readback_ms": parts[2] / steps * 1e3, "h2d_bytes_wire": int(st["h2d_bytes"])},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": fine_name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "issue_frac": issue, "algorithmic_bytes": int(alg_bytes), "kernel_ms": fine_ms, "kernel_ms_source": fine_src,
                     "peak_source": peak_src, "whole_step_frac": alg_bytes * reps / (ms_step * 1e-3) / 1e9 / peak},
        "stage_ms": stage, "stage_ms_note": "per flush, plain launches with events between stages: %.3f ms per flush" % (st_stages["ms_total"] / (steps * reps)),
    }
    if striped:
        rec["gather_ms"] = gather_step_ms
        if peer is not None:
            rec["delivery"] = {"how": "peer bands", "notify_ms": notify_ms, "render_only_ms": render_only_ms, "nccl_gather_ms": gather_step_ms}
            peer.close()
    if w == "c1":   # vkvg_surface_write_to_png timed separately (un-premultiply on the device, read-back, deflate on the host)
        path = "/tmp/vkvg_b200_bench_tiger_%d.png" % rank
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            assert surf.write_to_png(path) == 0
            ts.append(time.perf_counter() - t0)
        rec["write_to_png_ms"] = float(np.median(ts)) * 1e3
        try:
            os.remove(path)
        except OSError:
            pass
    ctx.close()
    surf.close()
    dev.close()
    return rec, frame


def slim(rec):
    """what a sub-record of the headline line keeps"""
    keep = ("metric", "value", "unit", "ms_per_step", "scaling", "stage_ms", "gather_ms", "delivery", "write_to_png_ms", "gpu_launches")
    out = {k: rec[k] for k in keep if k in rec}
    out["workload"] = rec["config"]["workload"]
    out["rule"] = rec["config"]["rule"]
    out["sharding"] = rec["config"]["sharding"]
    out["n_edges"], out["n_path_tiles"] = rec["config"]["n_edges"], rec["config"]["n_path_tiles"]
    out["roofline"] = {k: rec["roofline"][k] for k in ("kernel", "achieved", "peak", "frac", "algorithmic_bytes", "kernel_ms", "whole_step_frac")}
    out["e2e"] = {k: rec["e2e"][k] for k in ("value", "unit", "ms_per_step", "host_record_ms", "upload_render_ms", "readback_ms", "h2d_bytes_per_step", "d2h_bytes_per_step", "decoded_on")}
    return out


def run_ours(args):
    env = Env(args)
    w = args.workload
    steps, warmup = args.steps, max(args.warmup, 3)
    rec, frame = measure(env, w, args.rule, steps)
    line = {"metric": rec["metric"], "value": rec["value"], "unit": rec["unit"], "n_gpus": env.world, "steps": steps, "warmup": warmup,
            "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": rec["scaling"], "vs_baseline": None, "dtype": "f32+i64", "data": "synthetic"}
    line.update({k: rec[k] for k in ("config", "e2e", "gpu_launches", "clocks", "roofline", "stage_ms", "stage_ms_note") if k in rec})
    for k in ("gather_ms", "delivery", "write_to_png_ms"):
        if k in rec:
            line[k] = rec[k]
    sub_steps = max(3, min(steps, 10))
    if not args.only and w == "c2" and args.coverage == "msaa":
        # ---- the other single-GPU configurations, on rank 0's GPU (replicas of them measure nothing new) ----
        if env.rank == 0:
            cfgs = {}
            for name, (ww, rule) in (("c1", ("c1", "eo")), ("c2_eo", ("c2", "eo")), ("c3", ("c3", "nz")), ("c4", ("c4", "nz"))):
                solo = Env.__new__(Env)   # same device, no collective: only this rank takes part
                solo.__dict__.update(env.__dict__)
                solo.dist, solo.world = None, 1
                r, _ = measure(solo, ww, rule, sub_steps)
                cfgs[name] = slim(r)
            line["configs"] = cfgs
        env.barrier()
        # ---- the configurations that shard: strong scaling over the ranks of this run ----
        sh = {}
        for ww in ("c5a", "c5b"):
            r, _ = measure(env, ww, "nz", sub_steps, strong_world=env.world)
            sh[ww] = slim(r)
        line["sharded"] = sh
    if env.rank == 0 and env.world == 1 and not args.no_cpu_baseline:
        cb, px = cpu_baseline(w, args.rule)
        line["cpu_baseline"] = cb
        if px is not None and px.shape == frame.shape:
            from tests.parity import pixel_stats
            st = pixel_stats(frame, px)
            st["against"] = "the cpu_baseline leg's frame: " + REF_LABEL
            line["parity"] = st
    if env.rank == 0:
        print(json.dumps(line))
    if env.dist is not None:
        env.dist.destroy_process_group()


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
    ap.add_argument("--only", action="store_true", help="only the named workload: no configs / sharded sub-records")
    ap.add_argument("--no-graph", action="store_true", help="time plain kernel launches instead of CUDA graph replays")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

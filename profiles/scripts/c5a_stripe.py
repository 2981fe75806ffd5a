"""one rank's share of C5a on ONE GPU: stripe `rank` of `world` of the 16384^2 scene (what each of N ranks runs), for ncu / stage timing
    python profiles/scripts/c5a_stripe.py [rank world]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import bench, vkvg_b200 as v
from vkvg_b200 import sharding
rank, world = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3, 8)
dev = v.Device(4)
surf, y0, h = sharding.stripe_surface(dev, 16384, 16384, rank, world)
emit, units, info = bench.build_scene("c5a", 1, "nz")
cs = v.CommandStream(); emit(cs)
ctx = v.Context(surf)
assert ctx.submit(*cs.arrays2()) == 0
dev.set_profiling(True); dev.set_stage_timing(True)
dev.time_resident(surf, 2, True, True)
st = dev.time_resident(surf, 3, True, True)
print(json.dumps({"rank": rank, "world": world, "ms_total": st["ms_total"] / 3, "stage_ms": {k: x / 3 for k, x in st["ms_stage"].items()},
                  "n_edges": st["n_edges"], "n_tile_edges": st["n_tile_edges"], "n_path_tiles": st["n_nonempty"], "n_points": st["n_points"]}))

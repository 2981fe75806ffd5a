"""CPU: the parts of bench.py that run without a GPU — the synthetic scenes as packed command streams and the reference arm /
cpu_baseline leg (the reference's tessellation object code + the oracle's raster restatement on a bounded sample)."""
import numpy as np
import pytest

import bench
import vkvg_b200 as v


@pytest.mark.parametrize("workload,n_limit", [("c2", 30), ("c3", 500), ("c4", 10), ("c5a", 20), ("c1", 12)])
def test_scenes_pack_into_command_streams(workload, n_limit):
    emit, units, info = bench.build_scene(workload, 1, "nz", n_limit=n_limit)
    cs = v.CommandStream()
    emit(cs)
    ops, args = cs.arrays()
    assert ops.dtype == np.uint8 and args.dtype == np.float32 and len(ops) > 0 and len(args) > 0
    assert units > 0 and info["n_paths"] > 0 and info["n_segments"] > 0


def test_reference_worker_renders_a_bounded_sample(oracle_lib):
    dt, info, kind = bench._ref_worker(("c2", 1, "nz", 25, 0))
    assert dt > 0 and info["n_paths"] == 25
    assert kind == ("reference" if oracle_lib.ref_available() else "port")


def test_cpu_baseline_samples_are_defined_for_every_workload():
    for w in bench.SIZES:
        assert w in bench.SAMPLE and w in bench.CPU_SAMPLE and w in bench.FULL and w in bench.UNITS and w in bench.WORKLOAD_NAMES
        assert bench.build_units(w) > 0

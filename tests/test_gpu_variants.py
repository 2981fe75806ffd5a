"""GPU: the kernel variants a process can select - the first stroke emitter (float vertices, snap_verts_k, tri_edges_k), the new one with
and without its shared-memory staging area, the one-thread-per-element flatten counting pass - produce the SAME flattened points, stroke
vertices, indices, edges and pixels, bit for bit.  The switches are read once per process, so every variant renders in a process of its
own (tests/variant_worker.py) and reports digests."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ)
    for k in ("VKVG_B200_STROKE", "VKVG_B200_FLATTEN"):
        env.pop(k, None)
    env.update(env_extra)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "variant_worker.py")], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.fixture(scope="module")
def default_digests():
    return _run({})


def test_scene_is_not_trivial(default_digests):
    d = default_digests
    assert d["n_points"] > 2000
    assert d["dashed_round_n"][0] > 3 * 3000        # more than three vertices per item: blocks of the emitter go through the staging area
    assert d["wide_round_n"][0] > 12 * 300          # more than a block can stage: those write straight to global memory
    assert d["miter_n"][0] <= 3 * 1500
    assert all(d[k + "_n"][2] > 0 for k in ("dashed_round", "miter", "wide_round", "closed_round", "closed_dashed_bevel"))


@pytest.mark.parametrize("env", [{"VKVG_B200_STROKE": "legacy"}, {"VKVG_B200_STROKE": "direct"}, {"VKVG_B200_FLATTEN": "thread"}],
                         ids=["stroke_legacy", "stroke_direct", "flatten_thread"])
def test_variant_equals_default(default_digests, env):
    other = _run(env)
    diff = sorted(k for k in default_digests if default_digests[k] != other.get(k))
    assert not diff, diff

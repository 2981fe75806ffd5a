#!/bin/bash
# A/B of the stroke emitters (VKVG_B200_STROKE=legacy|direct|<unset>) and the per-frame launch count, on one GPU.
# usage: gpurun -- bash profiles/scripts/ab_stroke.sh     (logs in gpurun_out/)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t_new.log 2>&1; echo "pytest new rc=$?" | tee gpurun_out/t_rc.log
if ! grep -q " passed" gpurun_out/t_new.log || grep -q "failed" gpurun_out/t_new.log; then
  VKVG_B200_STROKE=legacy python -m pytest tests -m gpu -x -q > gpurun_out/t_legacy.log 2>&1; echo "pytest legacy rc=$?" | tee -a gpurun_out/t_rc.log
fi
tail -5 gpurun_out/t_new.log
for w in c1 c3; do
  python bench.py --only --workload $w --no-cpu-baseline > gpurun_out/b_${w}_new.json 2> gpurun_out/b_${w}_new.err
  VKVG_B200_STROKE=legacy python bench.py --only --workload $w --no-cpu-baseline > gpurun_out/b_${w}_legacy.json 2> gpurun_out/b_${w}_legacy.err
done
VKVG_B200_STROKE=direct python bench.py --only --workload c3 --no-cpu-baseline > gpurun_out/b_c3_direct.json 2> gpurun_out/b_c3_direct.err
python bench.py --only --workload c2 --no-cpu-baseline > gpurun_out/b_c2_new.json 2> gpurun_out/b_c2_new.err
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/b_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.4f ms" % d["ms_per_step"], "launches/step", d["gpu_launches"] / d["steps"], {k: round(v, 4) for k, v in d.get("stage_ms", {}).items()}, "e2e %.3f" % d["e2e"]["ms_per_step"])
    except Exception as e:
        print(f, "unreadable", e)
P

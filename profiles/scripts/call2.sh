#!/bin/bash
# full GPU suite, the three single-GPU workloads, then serialised launch lists (ncu, gpu__time_duration) of one frame each
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t_new.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/t_rc.log
tail -3 gpurun_out/t_new.log
for w in c1 c3 c2; do
  python bench.py --only --workload $w --no-cpu-baseline > gpurun_out/b_${w}_new.json 2> gpurun_out/b_${w}_new.err
done
VKVG_B200_STROKE=legacy python bench.py --only --workload c1 --no-cpu-baseline > gpurun_out/b_c1_legacy.json 2> gpurun_out/b_c1_legacy.err
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/b_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.4f ms" % d["ms_per_step"], "launches/step", d["gpu_launches"] / d["steps"], {k: round(v, 4) for k, v in d.get("stage_ms", {}).items()}, "e2e %.3f" % d["e2e"]["ms_per_step"])
    except Exception as e:
        print(f, "unreadable", e)
P
for w in c1 c2 c3; do
  timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches_${w}.csv python bench.py --only --workload $w --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${w}.log 2>&1
  echo "ncu $w rc=$?"
done

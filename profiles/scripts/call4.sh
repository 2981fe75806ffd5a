#!/bin/bash
# full GPU suite on the final tree, C5a / C3 alone, then one `ncu --set full` capture of the stroke / edge / binning kernels of a C3 frame
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t_new.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/t_rc.log
tail -3 gpurun_out/t_new.log
for w in c5a c3; do
  python bench.py --only --workload $w --no-cpu-baseline > gpurun_out/b_${w}_new.json 2> gpurun_out/b_${w}_new.err
done
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/b_c5a_new.json") + glob.glob("gpurun_out/b_c3_new.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "%.4f ms" % d["ms_per_step"], "launches/step", d["gpu_launches"] / d["steps"], {k: round(v, 4) for k, v in d.get("stage_ms", {}).items()}, "e2e %.3f" % d["e2e"]["ms_per_step"])
P
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'stroke_emit_k|tri_edges_warp_k|stroke_items_k|bin_count_k|bin_scatter_k|flatten_count_warp_k' -s 6 -c 5 -o gpurun_out/prof_stroke_c3_r2y python bench.py --only --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_c3.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep

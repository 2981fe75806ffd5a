#!/bin/bash
# full GPU suite, smoke(), then the default bench line (every configuration, cpu_baseline, parity record)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t_new.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/t_rc.log
tail -3 gpurun_out/t_new.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
( time timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real; echo "bench rc=$?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
def line(name, r):
    print("%-6s %9.4f ms  %12.1f %-14s fine %.4f  e2e %.3f ms  launches/step %.0f  %s" % (name, r["ms_per_step"], r["value"], r["unit"], r["roofline"]["kernel_ms"], r["e2e"]["ms_per_step"], r["gpu_launches"] / d["steps"], {k: round(v, 3) for k, v in r.get("stage_ms", {}).items()}))
line("c2", d)
for k, r in d.get("configs", {}).items(): line(k, r)
for k, r in d.get("sharded", {}).items(): line(k, r)
print("parity", d.get("parity"), "cpu_baseline", d.get("cpu_baseline", {}).get("value"))
P

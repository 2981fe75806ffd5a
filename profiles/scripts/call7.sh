#!/bin/bash
# serialised launch lists (ncu, gpu__time_duration) of C1 and C3 frames on the final tree
mkdir -p gpurun_out
for w in c1 c3; do
  timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches_${w}_r2z.csv python bench.py --only --workload $w --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${w}.log 2>&1
  echo "ncu $w rc=$?"
done

#!/bin/bash
# the full GPU suite (with tests/test_gpu_variants.py) on the final tree, and smoke()
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t_new.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/t_rc.log
tail -15 gpurun_out/t_new.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log

#!/bin/bash
# two GPUs: the IPC delivery test and the C5a stripes (strong scaling) on the final tree
mkdir -p gpurun_out
timeout 80 python -m pytest tests/test_gpu_delivery.py -m gpu -x -q 2>&1 | tail -3
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --only --workload c5a --no-cpu-baseline > gpurun_out/bench_c5a_n2_r2z.json 2> gpurun_out/bench_c5a_n2_r2z.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_c5a_n2_r2z.json

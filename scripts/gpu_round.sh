#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 1500 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err

#!/bin/bash
# all GPU tests (no -x) + a short bench line without the CPU / strong / latency legs
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu --durations=8 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== bench (short)"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-strong --no-latency ${BENCH_EXTRA:-} > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
cat gpurun_out/bench_quick.json; tail -5 gpurun_out/bench_quick.err

#!/bin/bash
# 1 GPU: whole suite with the native step driver (default) + host-resident window, default bench line
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
cat gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err

#!/bin/bash
# 1 GPU: window fast path (removal riding on particle_bcs, x-only classification) -- the window-related tests, then
# the whole suite, then the default bench
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "window or host_resident or exchange or protocol or loop_body or c1_deck" > gpurun_out/pytest_window.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_window.log
tail -12 gpurun_out/pytest_window.log
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
cat gpurun_out/bench_c3.json | cut -c1-1600; tail -3 gpurun_out/bench_c3.err
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log

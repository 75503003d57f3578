#!/bin/bash
# round 2, call 2: device-resident particle counts -- whole GPU suite, the default bench line (C3) with both exchange
# protocols, C2, reference arm
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
BENCH_EXCHANGE_CAPACITY=0 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c3_exact.json 2> gpurun_out/bench_c3_exact.err
python bench.py --steps 40 --warmup 3 --workload thermal_2048x256_m2_ppc64 --no-cpu > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
for f in bench_c3 bench_c3_exact bench_c2 bench_reference; do echo "== $f"; cat gpurun_out/$f.json; tail -3 gpurun_out/$f.err; done

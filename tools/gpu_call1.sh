#!/bin/bash
# round 2, first GPU call: whole GPU suite (no -x), the zz2 diagnostic, bench lines on C2 / C3 / C4 with and without
# deferred particle_bcs, the reference arm, and a launch list of the C3 bench.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=15 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
python tools/diag_zz2.py > gpurun_out/diag_zz2.txt 2>&1
cat gpurun_out/diag_zz2.txt
python bench.py --steps 20 --warmup 3 --workload lwfa_8192x512_m2_ppc32 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
BENCH_DEFERRED_BCS=1 python bench.py --steps 20 --warmup 3 --workload lwfa_8192x512_m2_ppc32 > gpurun_out/bench_c3_deferred.json 2> gpurun_out/bench_c3_deferred.err
python bench.py --steps 40 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
BENCH_DEFERRED_BCS=1 python bench.py --steps 40 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c2_deferred.json 2> gpurun_out/bench_c2_deferred.err
python bench.py --steps 10 --warmup 3 --workload modes5_4096x512_m5_ppc16 --no-e2e --no-cpu > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_c3.csv \
    python bench.py --steps 2 --warmup 3 --workload lwfa_8192x512_m2_ppc32 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
cat gpurun_out/bench_c3.json gpurun_out/bench_c3_deferred.json gpurun_out/bench_c2.json gpurun_out/bench_c2_deferred.json gpurun_out/bench_c4.json
tail -3 gpurun_out/bench_c3.err

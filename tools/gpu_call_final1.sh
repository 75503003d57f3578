#!/bin/bash
# 1 GPU, final build: the whole GPU suite, smoke(), the default bench (C3, with e2e and the CPU port), the reference
# arm in short, and the ncu evidence of the same command (launch list + one --set full capture of the push kernel).
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | cut -c1-300 | tee gpurun_out/r2g_pytest_gpu_n1.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2g_bench_c3_n1.json 2> gpurun_out/r2g_bench_c3_n1.err; tail -c 1500 gpurun_out/r2g_bench_c3_n1.json | cut -c1-1500
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2g_launches_c3.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_push_v2 -s 3 -c 1 -f -o gpurun_out/r2g_push_c3 \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_push.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3

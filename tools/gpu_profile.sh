#!/bin/bash
# ncu evidence of the default bench (C3): launch list (per-launch times: shares, not absolutes) and one --set full
# capture each of the fused push kernel and the two FDTD sweeps.  Numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_c3.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_push_v2 -s 3 -c 1 -f -o gpurun_out/r2_push_c3 \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_push.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_update_._bulk -s 6 -c 2 -f -o gpurun_out/r2_fields_c3 \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_fields.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -2 gpurun_out/ncu_push.log gpurun_out/ncu_fields.log

#!/bin/bash
# One gpurun call that settles everything written without a GPU in round 1's last session:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_first_call.sh'
# 1. the full GPU suite (the tests/test_zz*_gpu_*.py modules are the ones that have never run),
# 2. the default bench line and the reference arm,
# 3. a launch list of the bench under ncu (per-launch times: kernel shares, not absolutes).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
# the never-run modules on their own, without -x, so that one failure does not hide the others
python -m pytest tests/test_zz1_gpu_moments.py tests/test_zz2_gpu_counter_insert.py tests/test_zz3_gpu_sdf.py \
    tests/test_zz5_gpu_fullsize.py tests/test_zz4_gpu_gaussian_pulse.py tests/test_zz7_gpu_c1_hundred_steps.py \
    tests/test_zz6_gpu_deferred_bcs.py -m gpu -q --durations=15 > gpurun_out/pytest_zz.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_zz.log
tail -15 gpurun_out/pytest_zz.log
python bench.py --steps 40 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
BENCH_DEFERRED_BCS=1 python bench.py --steps 40 --warmup 3 > gpurun_out/bench_n1_deferred.json 2> gpurun_out/bench_n1_deferred.err
cat gpurun_out/bench_n1.json gpurun_out/bench_n1_deferred.json

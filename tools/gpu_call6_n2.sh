#!/bin/bash
# 2 GPUs: whole suite (NCCL tests included), default bench at N=1 and N=2 -- pre-sort on the side stream, merged window halos
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
cut -c1-1300 gpurun_out/bench_c3.json; tail -2 gpurun_out/bench_c3.err
BENCH_RANK_PHASES=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
   bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > gpurun_out/bench_c3_n2.json 2> gpurun_out/bench_c3_n2.err
cut -c1-1300 gpurun_out/bench_c3_n2.json; grep "^rank" gpurun_out/bench_c3_n2.err | sort -u; tail -2 gpurun_out/bench_c3_n2.err

#!/bin/bash
# 2 GPUs: whole suite (NCCL + mailbox tests included) with the side-stream particle chain, then the small LWFA bench at N=2
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log | cut -c1-400
CYLGPU_PRESORT=1 BENCH_RANK_PHASES=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
   bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --workload lwfa_1024x128_m2_ppc16 > gpurun_out/bench_small_n2.json 2> gpurun_out/bench_small_n2.err
cut -c1-200 gpurun_out/bench_small_n2.json; grep "^rank" gpurun_out/bench_small_n2.err | sort -u; tail -2 gpurun_out/bench_small_n2.err

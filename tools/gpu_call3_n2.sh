#!/bin/bash
# 2 GPUs: NCCL-transport parity tests, then the default bench (C3 z-decomposed over 2 GPUs) and the reference arm
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_nccl.py -m gpu -q > gpurun_out/pytest_nccl.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_nccl.log
tail -15 gpurun_out/pytest_nccl.log
for n in 2; do
  BENCH_RANK_PHASES=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
     bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_c3_n$n.json 2> gpurun_out/bench_c3_n$n.err
  echo "== N=$n"; cat gpurun_out/bench_c3_n$n.json; grep "^rank" gpurun_out/bench_c3_n$n.err; tail -2 gpurun_out/bench_c3_n$n.err
done

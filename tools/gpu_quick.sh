#!/bin/bash
# 2 GPUs, final binaries: the mailbox transport test and a short N=2 bench (barrier before the first exchange)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_nccl.py -m gpu -q -k "two_gpu_parity and mailboxes and not particles" 2>&1 | tail -2 | cut -c1-300 | tee gpurun_out/r2k_pytest_nccl_mailboxes.txt
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --workload lwfa_1024x128_m2_ppc16 2>&1 | tail -1 | cut -c1-400

#!/bin/bash
# 2 GPUs, final build: the transport tests (NCCL alone, callback, mailboxes for everything / for the particle messages)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nccl.py -m gpu -q 2>&1 | tail -3 | cut -c1-300 | tee gpurun_out/r2j_pytest_nccl_2gpu.txt

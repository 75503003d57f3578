#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_nccl.py -m gpu -q -k "four" 2>&1 | tail -12 | cut -c1-600 | tee gpurun_out/r2g_pytest_nccl_4gpu.txt

#!/bin/bash
# 1 GPU, final build: the boundary-line kernels after the reference-quirks switch (default path and switch off)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_zz4_gpu_gaussian_pulse.py -m gpu -q -x 2>&1 | tail -3 | cut -c1-300 | tee gpurun_out/r2k_pytest_quirks.txt

#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_zz8_gpu_rebalance.py -m gpu -q 2>&1 | grep -E "passed|failed|AssertionError: \(|^FAILED" | cut -c1-900

#!/bin/bash
# 1 GPU: the GPU test files on the top-hat and B-spline builds of the library against the oracle builds of the same
# shape (CYL_SHAPE selects both), and the default build's re-balancer file again
mkdir -p gpurun_out
{
for sh in tophat bspline3; do
  echo "== CYL_SHAPE=$sh"
  CYL_SHAPE=$sh timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz1_gpu_moments.py tests/test_zz2_gpu_counter_insert.py tests/test_zz4_gpu_gaussian_pulse.py tests/test_zz6_gpu_exchange_protocols.py tests/test_zz8_gpu_rebalance.py -m gpu -q 2>&1 | grep -E "^E   |passed|failed|^FAILED" | cut -c1-300 | head -30
done
echo "== triangle: re-balancer, variant 4"
timeout 600 python -m pytest tests/test_zz8_gpu_rebalance.py tests/test_gpu_parity.py -m gpu -q -k "rebalance or prescribed or 4 or variants_agree" 2>&1 | tail -2 | cut -c1-300
} 2>&1 | tee gpurun_out/r2h_shapes_gpu.txt

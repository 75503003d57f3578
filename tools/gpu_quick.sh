#!/bin/bash
mkdir -p gpurun_out
python tools/compare_math_variants.py 2>&1 | tail -8 | tee gpurun_out/r2g_math_variants.txt
CYLGPU_LIB=$PWD/cylindrical_epoch_b200/libcylgpu_refmath.so python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -2 | tee -a gpurun_out/r2g_math_variants.txt

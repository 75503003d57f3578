#!/bin/bash
# 1 GPU, last seconds of the budget: the laser / outflow decks on the B-spline and top-hat builds after the quirk switch went in
mkdir -p gpurun_out
for sh in bspline3 tophat; do CYL_SHAPE=$sh timeout 18 python -m pytest tests/test_gpu_parity.py tests/test_zz4_gpu_gaussian_pulse.py -m gpu -q -x -k "lwfa_steps or gaussian or quirks" 2>&1 | tail -1 | cut -c1-200; done | tee gpurun_out/r2k_pytest_shapes_after_quirks.txt

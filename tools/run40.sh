#!/bin/bash
# last 1.9 GPU-minutes of round 1: the cheapest informative probe of the kernels written without GPU access
mkdir -p gpurun_out
B200SEG_AUG_VERIFIED=1 B200SEG_VAT_VERIFIED=1 timeout -s KILL 100 python -m pytest tests/test_zz_gpu_aug.py tests/test_zz_gpu_vat.py -m gpu -q -x \
  -k "golden or affine_grid_sample or col2im or per_sample_norm" > gpurun_out/pytest_r40_probe.log 2>&1
echo "[pytest exit $?]" >> gpurun_out/pytest_r40_probe.log
tail -5 gpurun_out/pytest_r40_probe.log | cut -c1-200; grep -E "^E  *assert|^FAILED|Error" gpurun_out/pytest_r40_probe.log | head -10 | cut -c1-250

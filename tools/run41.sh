#!/bin/bash
# last GPU seconds of round 1: the EXACT (C = 19 / 21 / 2) instantiations of the aug kernel vs the oracle, ABI errors,
# then (if time remains) the input-gradient pass of DeepLab v2
mkdir -p gpurun_out
B200SEG_AUG_VERIFIED=1 B200SEG_VAT_VERIFIED=1 timeout -s KILL 25 python -m pytest tests/test_zz_gpu_aug.py tests/test_zz_gpu_vat.py -m gpu -q \
  -k "class_counts or rejects_bad or (input_gradient and deeplab_imagenet)" > gpurun_out/pytest_r41_probe.log 2>&1
echo "[pytest exit $?]" >> gpurun_out/pytest_r41_probe.log
tail -5 gpurun_out/pytest_r41_probe.log | cut -c1-200; grep -E "^E  *assert|^FAILED|Error" gpurun_out/pytest_r41_probe.log | head -10 | cut -c1-250

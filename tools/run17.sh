#!/bin/bash
mkdir -p gpurun_out
export B200SEG_SKIP_CPU_BASELINE=1
timeout -s KILL 900 python -m pytest tests/test_gpu_kernels.py -q -x -k "conv" > gpurun_out/pytest_conv.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_conv.log
B200SEG_SHAPE_PROFILE=gpurun_out/shape_profile7.txt timeout -s KILL 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench9.log 2>&1; echo "[bench exit $?]" >> gpurun_out/bench9.log
tail -3 gpurun_out/pytest_conv.log; grep -E "^E|^FAILED" gpurun_out/pytest_conv.log | head; tail -2 gpurun_out/bench9.log | cut -c1-300; head -14 gpurun_out/shape_profile7.txt

#!/bin/bash
# ICT kernels / iterations / entry points on the GPU, then the whole GPU suite
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_ict.py tests/test_gpu_entry_point.py -m gpu -q > gpurun_out/pytest_ict_r35.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_ict_r35.log
tail -25 gpurun_out/pytest_ict_r35.log | cut -c1-300
grep -E "^E  " gpurun_out/pytest_ict_r35.log | head -40 | cut -c1-300
timeout -s KILL 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_ict.py --deselect tests/test_gpu_entry_point.py > gpurun_out/pytest_gpu_r35.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_gpu_r35.log
tail -4 gpurun_out/pytest_gpu_r35.log

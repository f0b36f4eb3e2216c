#!/bin/bash
# final validation of HEAD: smoke(), then the whole GPU suite
mkdir -p gpurun_out
timeout -s KILL 120 python __graft_entry__.py smoke > gpurun_out/smoke_r38.log 2>&1; echo "[smoke exit $?]" >> gpurun_out/smoke_r38.log
tail -4 gpurun_out/smoke_r38.log | cut -c1-200
timeout -s KILL 400 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r38.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_gpu_r38.log
tail -4 gpurun_out/pytest_gpu_r38.log | cut -c1-200; grep -E "^E  *assert|^FAILED" gpurun_out/pytest_gpu_r38.log | head -8 | cut -c1-250

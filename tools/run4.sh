#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/bench1.log 2>&1; echo "[bench exit $?]" >> gpurun_out/bench1.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/bench1.log; tail -15 gpurun_out/pytest_gpu.log

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/aspp_bench.py 5 all > gpurun_out/aspp_bench2.log 2>&1; echo "[aspp exit $?]" >> gpurun_out/aspp_bench2.log
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/bench1.log 2>&1; echo "[bench exit $?]" >> gpurun_out/bench1.log
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_gpu.log
cat gpurun_out/aspp_bench2.log; tail -5 gpurun_out/bench1.log; tail -25 gpurun_out/pytest_gpu.log

#!/bin/bash
mkdir -p gpurun_out
export B200SEG_SKIP_CPU_BASELINE=1
timeout -s KILL 120 python tools/conv2_check.py > gpurun_out/conv2_check.log 2>&1; echo "[conv2 exit $?]" >> gpurun_out/conv2_check.log
timeout -s KILL 400 python tools/aspp_bench.py 5 all > gpurun_out/aspp_bench4.log 2>&1; echo "[aspp exit $?]" >> gpurun_out/aspp_bench4.log
timeout -s KILL 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_gpu.log
B200SEG_SHAPE_PROFILE=gpurun_out/shape_profile3.txt timeout -s KILL 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench4.log 2>&1; echo "[bench exit $?]" >> gpurun_out/bench4.log
cat gpurun_out/conv2_check.log; cat gpurun_out/aspp_bench4.log; tail -12 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/bench4.log | cut -c1-2200; head -16 gpurun_out/shape_profile3.txt

#!/bin/bash
# 2-CTA wgrad, separable NCHW bilinear backward, faster statistics reduction, 32-bit index math
mkdir -p gpurun_out
timeout -s KILL 300 python tools/aspp_bench.py 5 wg2 > gpurun_out/wg2.log 2>&1; echo "[wg2 exit $?]" >> gpurun_out/wg2.log
timeout -s KILL 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_gpu.log
timeout -s KILL 300 python tools/netops_bench.py 5 > gpurun_out/netops2.log 2>&1; echo "[netops exit $?]" >> gpurun_out/netops2.log
B200SEG_SKIP_CPU_BASELINE=1 B200SEG_SHAPE_PROFILE=gpurun_out/shape_profile_r22.txt timeout -s KILL 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r22.log 2>&1; echo "[bench exit $?]" >> gpurun_out/bench_r22.log
cat gpurun_out/wg2.log | tail -30
tail -3 gpurun_out/pytest_gpu.log; grep -E "^E |^FAILED|Error" gpurun_out/pytest_gpu.log | head -10
cat gpurun_out/netops2.log
tail -2 gpurun_out/bench_r22.log | cut -c1-400

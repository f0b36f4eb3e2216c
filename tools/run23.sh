#!/bin/bash
# session-3 validation of HEAD: GPU test-suite, default bench (with cpu_baseline) + shape profile, conv microbench
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x --durations=10 > gpurun_out/pytest_gpu.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_gpu.log
B200SEG_SHAPE_PROFILE=gpurun_out/shape_profile_r23.txt timeout -s KILL 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r23.log 2>&1; echo "[bench exit $?]" >> gpurun_out/bench_r23.log
timeout -s KILL 300 python tools/aspp_bench.py 5 all > gpurun_out/micro_r23.log 2>&1; echo "[micro exit $?]" >> gpurun_out/micro_r23.log
tail -14 gpurun_out/pytest_gpu.log; grep -E "^E |^FAILED|Error" gpurun_out/pytest_gpu.log | head -10
tail -2 gpurun_out/bench_r23.log | cut -c1-700

#!/bin/bash
# round-1 session-2 validation: full GPU test-suite, prefetch on/off microbench, default bench with shape profile
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_gpu.log
timeout -s KILL 420 python tools/aspp_bench.py 5 pf > gpurun_out/pf.log 2>&1; echo "[pf exit $?]" >> gpurun_out/pf.log
B200SEG_SHAPE_PROFILE=gpurun_out/shape_profile_r20.txt timeout -s KILL 900 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r20.log 2>&1; echo "[bench exit $?]" >> gpurun_out/bench_r20.log
tail -5 gpurun_out/pytest_gpu.log; grep -E "^E |^FAILED|Error" gpurun_out/pytest_gpu.log | head -20
cat gpurun_out/pf.log | tail -45
tail -2 gpurun_out/bench_r20.log | cut -c1-600

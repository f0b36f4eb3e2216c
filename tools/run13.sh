#!/bin/bash
mkdir -p gpurun_out
export B200SEG_SKIP_CPU_BASELINE=1
timeout -s KILL 120 python tools/conv2_check.py > gpurun_out/conv2_check.log 2>&1; echo "[conv2 exit $?]" >> gpurun_out/conv2_check.log
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "conv or wgrad or dgrad" > gpurun_out/pytest_conv.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_conv.log
timeout -s KILL 300 python tools/aspp_bench.py 1 trace > gpurun_out/trace2.log 2>&1; echo "[trace exit $?]" >> gpurun_out/trace2.log
timeout -s KILL 300 python tools/aspp_bench.py 5 epi > gpurun_out/epi_exp3.log 2>&1; echo "[epi exit $?]" >> gpurun_out/epi_exp3.log
timeout -s KILL 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench6.log 2>&1; echo "[bench exit $?]" >> gpurun_out/bench6.log
tail -3 gpurun_out/conv2_check.log; tail -4 gpurun_out/pytest_conv.log; grep -E "^---|tile \[|epi  tile|last" gpurun_out/trace2.log; cat gpurun_out/epi_exp3.log; tail -2 gpurun_out/bench6.log | cut -c1-1500

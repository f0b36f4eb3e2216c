#!/bin/bash
# pipeline timeline of the HBM-bound 1x1 layers after the warp-uniform change; DeepLab v2 bench line
mkdir -p gpurun_out
timeout -s KILL 120 python tools/aspp_bench.py 3 trace > gpurun_out/trace_r34.log 2>&1; echo "[trace exit $?]" >> gpurun_out/trace_r34.log
timeout -s KILL 120 python tools/aspp_bench.py 5 l3full > gpurun_out/l3full_r34.log 2>&1; echo "[l3full exit $?]" >> gpurun_out/l3full_r34.log
cat gpurun_out/l3full_r34.log
B200SEG_SKIP_CPU_BASELINE=1 timeout -s KILL 300 python bench.py --arch v2 --steps 10 --warmup 3 > gpurun_out/bench_r34_v2.log 2>&1; echo "[bench exit $?]" >> gpurun_out/bench_r34_v2.log
grep '^{' gpurun_out/bench_r34_v2.log | cut -c1-200

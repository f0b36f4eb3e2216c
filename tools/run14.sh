#!/bin/bash
mkdir -p gpurun_out
export B200SEG_SKIP_CPU_BASELINE=1
B200SEG_SHAPE_PROFILE=gpurun_out/shape_profile5.txt timeout -s KILL 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench7.log 2>&1; echo "[bench exit $?]" >> gpurun_out/bench7.log
tail -2 gpurun_out/bench7.log | cut -c1-400; head -45 gpurun_out/shape_profile5.txt

#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python tools/aspp_bench.py 1 trace > gpurun_out/trace.log 2>&1; echo "[trace exit $?]" >> gpurun_out/trace.log
cat gpurun_out/trace.log

#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python tools/aspp_bench.py 5 wg2 > gpurun_out/wg2.log 2>&1; echo "[wg2 exit $?]" >> gpurun_out/wg2.log
timeout -s KILL 300 python tools/aspp_bench.py 5 l3full > gpurun_out/l3full.log 2>&1; echo "[l3full exit $?]" >> gpurun_out/l3full.log
cat gpurun_out/wg2.log gpurun_out/l3full.log

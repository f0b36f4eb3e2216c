#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python tools/aspp_bench.py 5 l3full > gpurun_out/l3full.log 2>&1; echo "[exit $?]" >> gpurun_out/l3full.log
timeout -s KILL 600 ncu --set full --clock-control none -k regex:conv_gemm2 -c 8 -o gpurun_out/l3full python tools/aspp_bench.py 1 l3full > gpurun_out/ncu_l3full.log 2>&1; echo "[ncu exit $?]" >> gpurun_out/ncu_l3full.log
cat gpurun_out/l3full.log; tail -3 gpurun_out/ncu_l3full.log

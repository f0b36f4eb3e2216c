#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python tools/aspp_bench.py 5 epi > gpurun_out/epi_exp.log 2>&1; echo "[epi exit $?]" >> gpurun_out/epi_exp.log
timeout -s KILL 400 ncu --set full --import-source on --clock-control none -k regex:conv_gemm2 -c 1 -s 2 -o gpurun_out/l3_1x1_v5 python tools/aspp_bench.py 1 l3 > gpurun_out/ncu_l3.log 2>&1; echo "[ncu exit $?]" >> gpurun_out/ncu_l3.log
cat gpurun_out/epi_exp.log; tail -3 gpurun_out/ncu_l3.log

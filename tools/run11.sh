#!/bin/bash
mkdir -p gpurun_out
export B200SEG_SKIP_CPU_BASELINE=1
timeout -s KILL 400 python tools/aspp_bench.py 5 all > gpurun_out/aspp_bench5.log 2>&1; echo "[aspp exit $?]" >> gpurun_out/aspp_bench5.log
timeout -s KILL 300 python tools/aspp_bench.py 5 epi > gpurun_out/epi_exp2.log 2>&1; echo "[epi exit $?]" >> gpurun_out/epi_exp2.log
B200SEG_SHAPE_PROFILE=gpurun_out/shape_profile4.txt timeout -s KILL 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench5.log 2>&1; echo "[bench exit $?]" >> gpurun_out/bench5.log
cat gpurun_out/aspp_bench5.log gpurun_out/epi_exp2.log; tail -2 gpurun_out/bench5.log | cut -c1-3500; head -70 gpurun_out/shape_profile4.txt

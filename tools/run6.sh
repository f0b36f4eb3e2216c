#!/bin/bash
mkdir -p gpurun_out
B200SEG_SHAPE_PROFILE=gpurun_out/shape_profile.txt timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/bench2.log 2>&1; echo "[bench exit $?]" >> gpurun_out/bench2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 0 -c 1 -o gpurun_out/prof_conv_1x1 python tools/aspp_bench.py 1 l3 > gpurun_out/ncu3.log 2>&1
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/bench2.log; cat gpurun_out/shape_profile.txt; tail -12 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/ncu3.log

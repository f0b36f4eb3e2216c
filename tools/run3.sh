#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/net_probe.py > gpurun_out/net_probe.log 2>&1; echo "[net_probe exit $?]" >> gpurun_out/net_probe.log
timeout 600 python tools/aspp_bench.py 5 all > gpurun_out/aspp_bench.log 2>&1; echo "[aspp exit $?]" >> gpurun_out/aspp_bench.log
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/bench1.log 2>&1; echo "[bench exit $?]" >> gpurun_out/bench1.log
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 2 -c 2 -o gpurun_out/prof_conv_aspp python tools/aspp_bench.py 1 aspp > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_kernel -s 1 -c 1 -o gpurun_out/prof_wgrad_aspp python tools/aspp_bench.py 1 aspp > gpurun_out/ncu2.log 2>&1
tail -12 gpurun_out/net_probe.log; cat gpurun_out/aspp_bench.log; tail -5 gpurun_out/bench1.log; tail -15 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/ncu1.log

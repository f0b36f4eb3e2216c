#!/bin/bash
mkdir -p gpurun_out
export B200SEG_SKIP_CPU_BASELINE=1
timeout -s KILL 600 python bench.py --steps 6 --warmup 3 > gpurun_out/bench_graph.log 2>&1; echo "[bench exit $?]" >> gpurun_out/bench_graph.log
timeout -s KILL 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_gpu.log
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_eager.csv python bench.py --steps 1 --warmup 3 --eager --batch 16 > gpurun_out/ncu_launch.log 2>&1; echo "[ncu launches exit $?]" >> gpurun_out/ncu_launch.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:conv_gemm2_kernel -s 2 -c 1 -o gpurun_out/prof_conv2_aspp python tools/aspp_bench.py 1 aspp > gpurun_out/ncu4.log 2>&1
tail -2 gpurun_out/bench_graph.log | cut -c1-2500; tail -12 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/ncu_launch.log | cut -c1-300; wc -l gpurun_out/launches_eager.csv

#!/bin/bash
# PF default (K <= 512) + double-buffered single-operand prefetch; netops microbench; ncu launch list of one eager step
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_gpu.log
timeout -s KILL 300 python tools/aspp_bench.py 5 pf > gpurun_out/pf2.log 2>&1; echo "[pf exit $?]" >> gpurun_out/pf2.log
timeout -s KILL 300 python tools/netops_bench.py 5 > gpurun_out/netops.log 2>&1; echo "[netops exit $?]" >> gpurun_out/netops.log
B200SEG_SKIP_CPU_BASELINE=1 B200SEG_SHAPE_PROFILE=gpurun_out/shape_profile_r21.txt timeout -s KILL 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r21.log 2>&1; echo "[bench exit $?]" >> gpurun_out/bench_r21.log
B200SEG_SKIP_CPU_BASELINE=1 timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r21.csv python bench.py --eager --steps 1 --warmup 3 > gpurun_out/ncu_bench_r21.log 2>&1; echo "[ncu exit $?]" >> gpurun_out/ncu_bench_r21.log
gzip -f gpurun_out/launches_r21.csv
tail -3 gpurun_out/pytest_gpu.log; grep -E "^E |^FAILED|Error" gpurun_out/pytest_gpu.log | head -10
grep -E "pf=|identical" gpurun_out/pf2.log | head -50
cat gpurun_out/netops.log
tail -2 gpurun_out/bench_r21.log | cut -c1-400
tail -2 gpurun_out/ncu_bench_r21.log | cut -c1-300

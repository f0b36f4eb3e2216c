#!/bin/bash
# constant-divisor stem im2col + register-coefficient BN backward: operator parity; ICT kernel tests; bench (CutMix and ICT)
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_netops.py tests/test_gpu_ict.py -m gpu -q > gpurun_out/pytest_r36.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_r36.log
tail -4 gpurun_out/pytest_r36.log | cut -c1-200; grep -E "^E  *assert|^FAILED" gpurun_out/pytest_r36.log | head -10 | cut -c1-250
B200SEG_SKIP_CPU_BASELINE=1 timeout -s KILL 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r36.log 2>&1; echo "[bench exit $?]" >> gpurun_out/bench_r36.log
grep '^{' gpurun_out/bench_r36.log | cut -c1-160; tail -1 gpurun_out/bench_r36.log
B200SEG_SKIP_CPU_BASELINE=1 timeout -s KILL 400 python bench.py --loss ict --steps 10 --warmup 3 > gpurun_out/bench_r36_ict.log 2>&1; echo "[bench exit $?]" >> gpurun_out/bench_r36_ict.log
grep '^{' gpurun_out/bench_r36_ict.log | cut -c1-160; tail -1 gpurun_out/bench_r36_ict.log

#!/bin/bash
# warp-uniform (elect_one) producer / MMA-issuer code in all four conv kernels + division-free wgrad box walk:
# kernel parity first, then the pipeline-isolation microbench, the bench, and the whole GPU suite
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/pytest_kernels_r30.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_kernels_r30.log
tail -4 gpurun_out/pytest_kernels_r30.log
timeout -s KILL 200 python tools/aspp_bench.py 5 iso > gpurun_out/iso_r30.log 2>&1; echo "[iso exit $?]" >> gpurun_out/iso_r30.log
cat gpurun_out/iso_r30.log
B200SEG_SKIP_CPU_BASELINE=1 timeout -s KILL 400 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r30.log 2>&1; echo "[bench exit $?]" >> gpurun_out/bench_r30.log
grep '^{' gpurun_out/bench_r30.log | cut -c1-160; tail -1 gpurun_out/bench_r30.log
timeout -s KILL 600 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_kernels.py > gpurun_out/pytest_gpu_r30.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_gpu_r30.log
tail -5 gpurun_out/pytest_gpu_r30.log; grep -E "^E " gpurun_out/pytest_gpu_r30.log | head -10

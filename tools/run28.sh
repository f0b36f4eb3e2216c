#!/bin/bash
# fused Adam/SGD + EMA kernel, on-device evaluation kernel: GPU tests; bench with the fused optimiser vs torch's
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_gpu.log
B200SEG_SKIP_CPU_BASELINE=1 B200SEG_FUSED_OPT=1 timeout -s KILL 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r28_fused.log 2>&1; echo "[bench exit $?]" >> gpurun_out/bench_r28_fused.log
B200SEG_SKIP_CPU_BASELINE=1 B200SEG_FUSED_OPT=0 timeout -s KILL 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r28_torch.log 2>&1; echo "[bench exit $?]" >> gpurun_out/bench_r28_torch.log
B200SEG_SKIP_CPU_BASELINE=1 B200SEG_FUSED_OPT=1 timeout -s KILL 600 python bench.py --arch v2 --steps 8 --warmup 3 > gpurun_out/bench_r28_v2_fused.log 2>&1; echo "[bench exit $?]" >> gpurun_out/bench_r28_v2_fused.log
B200SEG_SKIP_CPU_BASELINE=1 B200SEG_FUSED_OPT=0 timeout -s KILL 600 python bench.py --arch v2 --steps 8 --warmup 3 > gpurun_out/bench_r28_v2_torch.log 2>&1; echo "[bench exit $?]" >> gpurun_out/bench_r28_v2_torch.log
tail -8 gpurun_out/pytest_gpu.log; grep -E "^E |^FAILED|Error" gpurun_out/pytest_gpu.log | head -20
for f in fused torch v2_fused v2_torch; do grep '^{' gpurun_out/bench_r28_$f.log | cut -c1-200; tail -1 gpurun_out/bench_r28_$f.log; done

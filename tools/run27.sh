#!/bin/bash
# batched-trunk iteration + softmax reciprocal: GPU tests, bench (batched vs pass-by-pass), HBM-op ncu captures
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_gpu.log
B200SEG_SKIP_CPU_BASELINE=1 B200SEG_SHAPE_PROFILE=gpurun_out/shape_profile_r27.txt timeout -s KILL 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r27.log 2>&1; echo "[bench exit $?]" >> gpurun_out/bench_r27.log
B200SEG_SKIP_CPU_BASELINE=1 B200SEG_BATCH_TRUNK=0 timeout -s KILL 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r27_nobatch.log 2>&1; echo "[bench exit $?]" >> gpurun_out/bench_r27_nobatch.log
timeout -s KILL 200 python tools/netops_bench.py 5 consistency > gpurun_out/netops_r27.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none -k regex:'consistency_kernel|ce_kernel|ema_|mix_kernel|bn_apply|bn_stats|bn_bwd|bilinear|maxpool|im2col|relu_gate|gap_' -c 40 -o gpurun_out/hbm_ops_r27 -f python tools/netops_bench.py 0 > gpurun_out/ncu_hbm_r27.log 2>&1; echo "[ncu hbm exit $?]" >> gpurun_out/ncu_hbm_r27.log
python tools/ncu_summary.py gpurun_out/hbm_ops_r27.ncu-rep > gpurun_out/hbm_ops_r27_summary.txt 2>&1
tail -8 gpurun_out/pytest_gpu.log; grep -E "^E |^FAILED|Error" gpurun_out/pytest_gpu.log | head -10
grep '^{' gpurun_out/bench_r27.log | cut -c1-300; grep '^{' gpurun_out/bench_r27_nobatch.log | cut -c1-300; tail -3 gpurun_out/bench_r27.log | cut -c1-300
cat gpurun_out/netops_r27.log; tail -2 gpurun_out/ncu_hbm_r27.log

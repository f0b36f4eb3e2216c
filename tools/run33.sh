#!/bin/bash
# 2-GPU data-parallel run of the bench (torchrun, NCCL all-reduce of the flat gradient buffer, fused optimiser + EMA)
mkdir -p gpurun_out
B200SEG_SKIP_CPU_BASELINE=1 timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_r33_2gpu.log 2>&1; echo "[bench exit $?]" >> gpurun_out/bench_r33_2gpu.log
grep '^{' gpurun_out/bench_r33_2gpu.log | cut -c1-200; tail -2 gpurun_out/bench_r33_2gpu.log | cut -c1-300
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_r33_ref2.log 2>&1; echo "[ref exit $?]" >> gpurun_out/bench_r33_ref2.log
tail -2 gpurun_out/bench_r33_ref2.log | cut -c1-400

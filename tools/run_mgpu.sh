#!/bin/bash
# usage: tools/run_mgpu.sh N   (inside gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
export B200SEG_SKIP_CPU_BASELINE=1
nvidia-smi -L > gpurun_out/mgpu_${N}.log 2>&1
timeout -s KILL 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 6 --warmup 3 >> gpurun_out/mgpu_${N}.log 2>&1; echo "[bench exit $?]" >> gpurun_out/mgpu_${N}.log
grep -E "^\{|Error|error|exit" gpurun_out/mgpu_${N}.log | cut -c1-900 | tail -8

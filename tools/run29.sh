#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_optim.py -m gpu -q > gpurun_out/pytest_optim.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_optim.log
timeout -s KILL 300 python tools/aspp_bench.py 5 major > gpurun_out/major.log 2>&1; echo "[major exit $?]" >> gpurun_out/major.log
tail -5 gpurun_out/pytest_optim.log; grep -E "^E " gpurun_out/pytest_optim.log | head; cat gpurun_out/major.log

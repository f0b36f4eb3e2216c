#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python tools/aspp_bench.py 5 iso > gpurun_out/iso.log 2>&1; echo "[iso exit $?]" >> gpurun_out/iso.log
cat gpurun_out/iso.log

#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python tools/gpu_probe.py netops 2>&1 | grep -v " OK$" ) > gpurun_out/netops.log 2>&1
echo "[netops exit $?]" >> gpurun_out/netops.log
timeout 900 python tools/net_probe.py > gpurun_out/net_probe.log 2>&1
echo "[net_probe exit $?]" >> gpurun_out/net_probe.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "[smoke exit $?]" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench1.log 2>&1
echo "[bench exit $?]" >> gpurun_out/bench1.log
tail -30 gpurun_out/netops.log; tail -30 gpurun_out/net_probe.log; tail -15 gpurun_out/smoke.log; tail -12 gpurun_out/bench1.log

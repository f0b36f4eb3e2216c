#!/bin/bash
# Runs each probe group in its own process with a timeout so one hang cannot block the rest.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/probe.log 2>&1
for g in "$@"; do
  if [ "$g" == "wdebug" ]; then
    timeout 200 python tools/wgrad_debug.py >> gpurun_out/probe.log 2>&1
  else
    timeout 240 python tools/gpu_probe.py $g >> gpurun_out/probe.log 2>&1
  fi
  echo "[group $g exit $?]" >> gpurun_out/probe.log
done
grep -v " OK$" gpurun_out/probe.log | tail -150

#!/bin/bash
# prefetch equivalence, im2col grid change, bench with prefetch in the e2e loop
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_graph.py tests/test_gpu_netops.py -m gpu -q > gpurun_out/pytest_r37.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_r37.log
tail -4 gpurun_out/pytest_r37.log | cut -c1-200; grep -E "^E  *assert|^FAILED|Error" gpurun_out/pytest_r37.log | head -10 | cut -c1-250
B200SEG_SKIP_CPU_BASELINE=1 timeout -s KILL 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r37.log 2>&1; echo "[bench exit $?]" >> gpurun_out/bench_r37.log
grep '^{' gpurun_out/bench_r37.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print(d['value'], d['ms_per_step'], d['clocks'], 'e2e', d['e2e'])
print({k:v['ms'] for k,v in list(r['per_kernel'].items())[:10]})
"; tail -1 gpurun_out/bench_r37.log

#!/bin/bash
# rerun of the parts of run31 that hit the int16 box-interval limit: wgrad parity, full bench line, ncu launch list
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_nets.py -m gpu -q -x > gpurun_out/pytest_kernels_r32.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_kernels_r32.log
tail -3 gpurun_out/pytest_kernels_r32.log
B200SEG_SHAPE_PROFILE=gpurun_out/shape_profile_r32.txt timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r32.log 2>&1; echo "[bench exit $?]" >> gpurun_out/bench_r32.log
grep '^{' gpurun_out/bench_r32.log | cut -c1-160; tail -1 gpurun_out/bench_r32.log
timeout -s KILL 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/launches_r32.csv python tools/ncu_step.py > gpurun_out/ncu_step_r32.log 2>&1; echo "[ncu step exit $?]" >> gpurun_out/ncu_step_r32.log
tail -2 gpurun_out/ncu_step_r32.log; wc -l gpurun_out/launches_r32.csv
python tools/launch_list_summary.py gpurun_out/launches_r32.csv > gpurun_out/launch_list_summary_r32.txt 2>&1; head -12 gpurun_out/launch_list_summary_r32.txt

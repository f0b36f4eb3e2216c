#!/bin/bash
# evidence run: HBM-bound operators (event timings + ncu full captures), launch list + DRAM traffic of one eager step
mkdir -p gpurun_out
timeout -s KILL 300 python tools/netops_bench.py 5 > gpurun_out/netops_r26.log 2>&1; echo "[netops exit $?]" >> gpurun_out/netops_r26.log
timeout -s KILL 600 ncu --set full --clock-control none -k regex:'consistency|ce_kernel|ema_|mix_kernel|bn_apply|bn_stats|bn_bwd|bilinear|maxpool|im2col' -c 40 -o gpurun_out/hbm_ops_r26 -f python tools/netops_bench.py 0 > gpurun_out/ncu_hbm_r26.log 2>&1; echo "[ncu hbm exit $?]" >> gpurun_out/ncu_hbm_r26.log
python tools/ncu_summary.py gpurun_out/hbm_ops_r26.ncu-rep > gpurun_out/hbm_ops_r26_summary.txt 2>&1
timeout -s KILL 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/launches_r26.csv python tools/ncu_step.py > gpurun_out/ncu_step_r26.log 2>&1; echo "[ncu step exit $?]" >> gpurun_out/ncu_step_r26.log
cat gpurun_out/netops_r26.log; tail -3 gpurun_out/ncu_hbm_r26.log; tail -3 gpurun_out/ncu_step_r26.log; wc -l gpurun_out/launches_r26.csv; ls -la gpurun_out

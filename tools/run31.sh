#!/bin/bash
# closed-form stage count in the wgrad issuer + evidence for the warp-uniform conv kernels:
# kernel parity, pipeline isolation, full bench line (with the CPU baseline), ncu launch list of one eager step,
# ncu --set full of the ASPP fprop / dgrad / wgrad kernels and of the HBM-bound layer3 1x1
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/pytest_kernels_r31.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_kernels_r31.log
tail -3 gpurun_out/pytest_kernels_r31.log
timeout -s KILL 200 python tools/aspp_bench.py 5 iso > gpurun_out/iso_r31.log 2>&1; echo "[iso exit $?]" >> gpurun_out/iso_r31.log
grep wgrad gpurun_out/iso_r31.log
timeout -s KILL 200 python tools/aspp_bench.py 5 all > gpurun_out/micro_r31.log 2>&1; echo "[micro exit $?]" >> gpurun_out/micro_r31.log
B200SEG_SHAPE_PROFILE=gpurun_out/shape_profile_r31.txt timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r31.log 2>&1; echo "[bench exit $?]" >> gpurun_out/bench_r31.log
grep '^{' gpurun_out/bench_r31.log | cut -c1-160; tail -1 gpurun_out/bench_r31.log
timeout -s KILL 600 ncu --set full --import-source on --clock-control none -k regex:'conv_gemm2|conv_wgrad2' -c 6 -o gpurun_out/aspp_r31 -f python tools/aspp_bench.py 1 aspp > gpurun_out/ncu_aspp_r31.log 2>&1; echo "[ncu aspp exit $?]" >> gpurun_out/ncu_aspp_r31.log
python tools/ncu_summary.py gpurun_out/aspp_r31.ncu-rep > gpurun_out/aspp_r31_summary.txt 2>&1
timeout -s KILL 300 ncu --set full --clock-control none -k regex:'conv_gemm2|conv_wgrad2' -c 4 -o gpurun_out/l3_r31 -f python tools/aspp_bench.py 1 l3 > gpurun_out/ncu_l3_r31.log 2>&1; echo "[ncu l3 exit $?]" >> gpurun_out/ncu_l3_r31.log
python tools/ncu_summary.py gpurun_out/l3_r31.ncu-rep > gpurun_out/l3_r31_summary.txt 2>&1
timeout -s KILL 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/launches_r31.csv python tools/ncu_step.py > gpurun_out/ncu_step_r31.log 2>&1; echo "[ncu step exit $?]" >> gpurun_out/ncu_step_r31.log
tail -2 gpurun_out/ncu_aspp_r31.log; tail -2 gpurun_out/ncu_step_r31.log; wc -l gpurun_out/launches_r31.csv
grep -E "tensor_cycles_active_realtime|time_duration|kernel:" gpurun_out/aspp_r31_summary.txt | head -30

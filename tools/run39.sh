#!/bin/bash
# FIRST GPU call of the next session: validates everything written after the round-1 GPU budget was spent
# (augmentation-consistency kernels, VAT kernels + input-gradient pass), then confirms the headline path is unchanged.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/run39.sh'
mkdir -p gpurun_out
# 1. the binding form of the two pending test files (non-strict xfail marker off)
B200SEG_AUG_VERIFIED=1 B200SEG_VAT_VERIFIED=1 B200SEG_DL3_VERIFIED=1 B200SEG_UNET_VERIFIED=1 timeout -s KILL 900 python -m pytest tests/test_zz_gpu_aug.py tests/test_zz_gpu_vat.py tests/test_zzy_gpu_deeplab3.py tests/test_zzy_gpu_resunet.py tests/test_zzy_gpu_denseunet.py tests/test_zzz_gpu_input.py -m gpu -q \
  > gpurun_out/pytest_r39_aug_vat.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_r39_aug_vat.log
tail -4 gpurun_out/pytest_r39_aug_vat.log | cut -c1-200; grep -E "^E  *assert|^FAILED|Error" gpurun_out/pytest_r39_aug_vat.log | head -20 | cut -c1-250
# 2. loss-kernel regression (the per-pixel loss tail was factored out of consistency_kernel; SASS instruction mix unchanged)
timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_ict.py -m gpu -q > gpurun_out/pytest_r39_losses.log 2>&1
echo "[pytest exit $?]" >> gpurun_out/pytest_r39_losses.log; tail -3 gpurun_out/pytest_r39_losses.log | cut -c1-200
# 3. bench lines: headline (unchanged path), aug, vat
for loss in cutmix aug vat; do
  B200SEG_SKIP_CPU_BASELINE=1 timeout -s KILL 400 python bench.py --steps 10 --warmup 3 --loss $loss > gpurun_out/bench_r39_$loss.log 2>&1
  echo "[bench exit $?]" >> gpurun_out/bench_r39_$loss.log
  grep '^{' gpurun_out/bench_r39_$loss.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$loss', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'])" 2>/dev/null || tail -3 gpurun_out/bench_r39_$loss.log | cut -c1-300
done
# 3b. BASELINE config 4: DenseNet-161 U-Net, 224x224, augmentation consistency
B200SEG_SKIP_CPU_BASELINE=1 timeout -s KILL 400 python bench.py --steps 10 --warmup 3 --arch denseunet --loss aug > gpurun_out/bench_r39_config4.log 2>&1
echo "[bench exit $?]" >> gpurun_out/bench_r39_config4.log; tail -2 gpurun_out/bench_r39_config4.log | cut -c1-400
# 3c. uint8 batches across PCIe (extra key e2e_u8)
B200SEG_SKIP_CPU_BASELINE=1 timeout -s KILL 400 python bench.py --steps 10 --warmup 3 --u8-inputs > gpurun_out/bench_r39_u8.log 2>&1; tail -1 gpurun_out/bench_r39_u8.log | cut -c1-200
# 4. one ncu --set full capture of the new loss kernel (tests drive it at C = 19 / 21)
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:aug_consistency_kernel -c 2 \
  -o gpurun_out/aug_r39 python -m pytest tests/test_zz_gpu_aug.py -m gpu -q -k "class_counts and var" > gpurun_out/ncu_aug_r39.log 2>&1
ls -la gpurun_out/aug_r39.ncu-rep 2>/dev/null

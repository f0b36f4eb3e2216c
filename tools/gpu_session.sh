#!/bin/bash
# One parametrised launcher for the GPU calls of a round (replaces the per-call scratch scripts of round 1):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_session.sh <stage> [tag]'
# Every stage writes its logs under gpurun_out/ with the given tag; nothing here is a bench value unless bench.py printed it
# outside a profiler.
stage=${1:-tests}; tag=${2:-r02}
mkdir -p gpurun_out
run_tests() {   # $1 = log tag, rest = pytest args
  local t=$1; shift
  B200SEG_PARITY_LOG=gpurun_out/parity_$t.txt timeout -s KILL 1500 python -m pytest "$@" -m gpu -q --tb=short -p no:cacheprovider \
    > gpurun_out/pytest_$t.log 2>&1
  echo "[pytest exit $?]" >> gpurun_out/pytest_$t.log
  tail -5 gpurun_out/pytest_$t.log | cut -c1-220
  grep -E "^(FAILED|ERROR)" gpurun_out/pytest_$t.log | head -40 | cut -c1-220
}
bench_line() {  # $1 = log name, rest = bench args
  local n=$1; shift
  timeout -s KILL 900 python bench.py "$@" > gpurun_out/bench_$n.log 2>&1
  echo "[bench exit $?]" >> gpurun_out/bench_$n.log
  grep '^{' gpurun_out/bench_$n.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('$n', d.get('value'), 'img/s', d.get('ms_per_step'), 'ms e2e', d.get('e2e',{}).get('value'), d.get('clocks'), 'modes', d.get('precision_modes'))
    for k in ('parity','incumbent_cudnn'):
        if k in d: print(' ', k, json.dumps(d[k])[:1500])
    r=d.get('roofline',{}); print('  roofline', r.get('kernel'), r.get('achieved'), r.get('frac'), r.get('share_of_step'), r.get('tf32_matmul_measured'))
" 2>/dev/null || tail -5 gpurun_out/bench_$n.log | cut -c1-300
}
case $stage in
  tests)      run_tests $tag tests ;;
  first)      # first call of the round: whole suite (binding), producer-order A/B, headline bench with parity + incumbent
    run_tests $tag tests
    for v in 0 1; do B200SEG_TAP_OUTER=$v timeout -s KILL 300 python tools/aspp_bench.py 3 all > gpurun_out/micro_${tag}_tapouter$v.log 2>&1; done
    paste -d'|' <(cut -c1-75 gpurun_out/micro_${tag}_tapouter0.log) <(cut -c46-75 gpurun_out/micro_${tag}_tapouter1.log) | head -40
    B200SEG_SHAPE_PROFILE=gpurun_out/shape_profile_$tag.txt bench_line $tag --steps 10 --warmup 3
    ;;
  second)     # TMA epilogue A/B + bit-exactness, the tests that failed in the first call, a bench line
    timeout -s KILL 600 python tools/aspp_bench.py 3 tma > gpurun_out/micro_${tag}_tma.log 2>&1; cut -c1-200 gpurun_out/micro_${tag}_tma.log
    run_tests $tag tests/test_gpu_kernels.py tests/test_gpu_fullsize.py tests/test_zzy_gpu_deeplab3.py tests/test_zzy_gpu_denseunet.py tests/test_gpu_nets.py tests/test_gpu_netops.py
    grep -h "denseunet\|dl3 (v3)" gpurun_out/pytest_$tag.log | head; cat gpurun_out/parity_$tag.txt | grep -v "^conv cfg3" | cut -c1-250
    B200SEG_SKIP_EXTRAS=1 B200SEG_SKIP_CPU_BASELINE=1 B200SEG_SHAPE_PROFILE=gpurun_out/shape_profile_$tag.txt bench_line $tag --steps 10 --warmup 3 --no-second-precision --no-tf32-peak
    head -30 gpurun_out/shape_profile_$tag.txt
    ;;
  third)      # multi-tensor prefills + parity-mode K chunking + loop variants: whole suite; bench; ncu --set full of the ASPP kernels
    run_tests $tag tests
    grep -h "denseunet\|dl3 (v3)\|pi / cutout\|per_pixel" gpurun_out/pytest_$tag.log | head -20; grep "fullsize\|cutmix iter" gpurun_out/parity_$tag.txt | grep -v " tf32 " | cut -c1-230
    B200SEG_SKIP_EXTRAS=1 B200SEG_SKIP_CPU_BASELINE=1 B200SEG_SHAPE_PROFILE=gpurun_out/shape_profile_$tag.txt bench_line $tag --steps 10 --warmup 3 --no-second-precision --no-tf32-peak
    timeout -s KILL 600 ncu --set full --import-source on --clock-control none -k regex:'conv_gemm2|conv_wgrad2' -c 6 -o gpurun_out/aspp_$tag -f python tools/aspp_bench.py 1 aspp > gpurun_out/ncu_aspp_$tag.log 2>&1; echo "[ncu aspp exit $?]" >> gpurun_out/ncu_aspp_$tag.log
    python tools/ncu_summary.py gpurun_out/aspp_$tag.ncu-rep > gpurun_out/aspp_${tag}_summary.txt 2>&1
    grep -E "tensor_cycles_active_realtime|time_duration|kernel:|dram__bytes|lts__t_bytes|lts__throughput" gpurun_out/aspp_${tag}_summary.txt | head -60
    ;;
  fourth)     # whole suite + bench (no extras)
    run_tests $tag tests
    grep -h "denseunet\|dl3 (v3)\|pi / cutout\|per_pixel" gpurun_out/pytest_$tag.log | head -20; grep "fullsize\|cutmix iter" gpurun_out/parity_$tag.txt | grep -v " tf32 " | cut -c1-230
    B200SEG_SKIP_EXTRAS=1 B200SEG_SKIP_CPU_BASELINE=1 B200SEG_SHAPE_PROFILE=gpurun_out/shape_profile_$tag.txt bench_line $tag --steps 10 --warmup 3 --no-second-precision --no-tf32-peak
    head -12 gpurun_out/shape_profile_$tag.txt
    ;;
  ddp2)       # 2 GPUs: correctness of the bucketed / overlapped all-reduce, then bench A/B (buckets 4 vs 1)
    timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py > gpurun_out/ddp_check_$tag.log 2>&1
    grep -E "buckets=|overlapped|DDP_CHECK|Error|error" gpurun_out/ddp_check_$tag.log | head -20
    for b in 4 1; do
      B200SEG_GRAD_BUCKETS=$b B200SEG_SKIP_CPU_BASELINE=1 timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_${tag}_2gpu_buckets$b.log 2>&1
      grep '^{' gpurun_out/bench_${tag}_2gpu_buckets$b.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('buckets $b', d['value'], 'img/s', d['ms_per_step'], 'ms e2e', d['e2e']['value'], d['clocks'])" || tail -5 gpurun_out/bench_${tag}_2gpu_buckets$b.log
    done
    ;;
  evidence)   # bench lines of the sibling loops / BASELINE config 4, ncu --set full of the aug-consistency kernel, ncu launch list of one step
    for spec in "aug:--loss aug" "vat:--loss vat" "ict:--loss ict" "config4:--arch denseunet --loss aug" "cfg2:--arch v2"; do
      n=${spec%%:*}; a=${spec#*:}
      B200SEG_SKIP_EXTRAS=1 B200SEG_SKIP_CPU_BASELINE=1 bench_line ${tag}_$n --steps 10 --warmup 3 --no-second-precision --no-tf32-peak $a
    done
    timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:aug_consistency_kernel -c 2 -o gpurun_out/aug_$tag -f python -m pytest tests/test_zz_gpu_aug.py -m gpu -q -k "class_counts and var" > gpurun_out/ncu_aug_$tag.log 2>&1
    python tools/ncu_summary.py gpurun_out/aug_$tag.ncu-rep > gpurun_out/aug_${tag}_summary.txt 2>&1; grep -E "kernel:|time_duration|dram__bytes|dram_throughput" gpurun_out/aug_${tag}_summary.txt | head -12
    timeout -s KILL 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/launches_$tag.csv python tools/ncu_step.py > gpurun_out/ncu_step_$tag.log 2>&1; echo "[ncu step exit $?]" >> gpurun_out/ncu_step_$tag.log
    python tools/launch_list_summary.py gpurun_out/launches_$tag.csv > gpurun_out/launch_list_summary_$tag.txt 2>&1; head -25 gpurun_out/launch_list_summary_$tag.txt
    gzip -f gpurun_out/launches_$tag.csv
    ;;
  fifth)      # colour-jitter / input kernels + whole suite, pipeline isolation of the conv kernels, ncu of the aug loss kernel at full size
    run_tests $tag tests
    timeout -s KILL 300 python tools/aspp_bench.py 3 iso > gpurun_out/iso_$tag.log 2>&1; cut -c1-110 gpurun_out/iso_$tag.log
    B200SEG_SKIP_EXTRAS=1 B200SEG_SKIP_CPU_BASELINE=1 timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:aug_consistency_kernel -c 2 -o gpurun_out/aug_$tag -f python bench.py --steps 1 --warmup 3 --loss aug --eager --no-second-precision --no-tf32-peak > gpurun_out/ncu_aug_$tag.log 2>&1
    python tools/ncu_summary.py gpurun_out/aug_$tag.ncu-rep > gpurun_out/aug_${tag}_summary.txt 2>&1; grep -E "kernel:|time_duration|dram__bytes|dram_throughput|registers" gpurun_out/aug_${tag}_summary.txt | head -12
    ;;
  timeline)   # device timeline of the graph-replayed iteration (CUPTI via torch.profiler) + the vendor's TF32 GEMM under ncu
    timeout -s KILL 600 python tools/timeline.py v3plus gpurun_out/timeline_$tag.txt > gpurun_out/timeline_$tag.log 2>&1; head -75 gpurun_out/timeline_$tag.txt | cut -c1-200 || tail -20 gpurun_out/timeline_$tag.log
    timeout -s KILL 300 python tools/cublas_tf32_probe.py --time > gpurun_out/cublas_tf32_$tag.log 2>&1; cat gpurun_out/cublas_tf32_$tag.log | tail -4
    timeout -s KILL 400 ncu --set full --clock-control none --profile-from-start off -o gpurun_out/cublas_tf32_$tag -f python tools/cublas_tf32_probe.py > gpurun_out/ncu_cublas_$tag.log 2>&1
    python tools/ncu_summary.py gpurun_out/cublas_tf32_$tag.ncu-rep > gpurun_out/cublas_tf32_${tag}_summary.txt 2>&1; grep -E "kernel:|time_duration|tensor_cycles|mem_tensor|dram__bytes|lts__throughput|shared_mem_per_block|grid_size|block_size" gpurun_out/cublas_tf32_${tag}_summary.txt | head -40
    ;;
  stages)     # 5 vs 6 operand stages of the compute-bound CTA-pair kernel (knob 10), kernel tests with the swizzled staging tile, pipeline trace
    timeout -s KILL 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_$tag.log 2>&1; tail -3 gpurun_out/pytest_$tag.log | cut -c1-200
    for v in 5 6; do B200SEG_MAIN_STAGES=$v timeout -s KILL 300 python tools/aspp_bench.py 5 all > gpurun_out/micro_${tag}_stages$v.log 2>&1; done
    paste -d'|' <(cut -c1-78 gpurun_out/micro_${tag}_stages5.log) <(cut -c46-78 gpurun_out/micro_${tag}_stages6.log) | head -40
    timeout -s KILL 300 python tools/aspp_bench.py 3 trace3 > gpurun_out/trace3_$tag.log 2>&1; cat gpurun_out/trace3_$tag.log | cut -c1-250
    ;;
  headab)     # whole suite on HEAD (batched head default), then bench A/B: head over all mini-batches vs head per mini-batch
    run_tests $tag tests
    grep "fullsize\|cutmix iter" gpurun_out/parity_$tag.txt | grep -v " tf32 " | cut -c1-230
    for v in 1 0; do
      B200SEG_BATCH_HEAD=$v B200SEG_SKIP_EXTRAS=1 B200SEG_SKIP_CPU_BASELINE=1 B200SEG_SHAPE_PROFILE=gpurun_out/shape_profile_${tag}_head$v.txt bench_line ${tag}_head$v --steps 10 --warmup 3 --no-second-precision --no-tf32-peak
    done
    B200SEG_PDL=1 B200SEG_SKIP_EXTRAS=1 B200SEG_SKIP_CPU_BASELINE=1 bench_line ${tag}_pdl1 --steps 10 --warmup 3 --no-second-precision --no-tf32-peak
    B200SEG_PDL=1 timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_graph.py tests/test_gpu_nets.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_${tag}_pdl1.log 2>&1; tail -3 gpurun_out/pytest_${tag}_pdl1.log | cut -c1-200
    head -45 gpurun_out/shape_profile_${tag}_head1.txt
    ;;
  aspp32)     # ncu --set full of the ASPP 3x3 d12 kernels and of the layer3 3x3 at the launch shapes of the batched-head iteration (32 images)
    timeout -s KILL 300 python tools/aspp_bench.py 5 aspp32 > gpurun_out/micro_${tag}_aspp32.log 2>&1; timeout -s KILL 300 python tools/aspp_bench.py 5 l3x3 >> gpurun_out/micro_${tag}_aspp32.log 2>&1; cat gpurun_out/micro_${tag}_aspp32.log | cut -c1-120
    for w in aspp32 l3x3; do
      timeout -s KILL 600 ncu --set full --import-source on --clock-control none -k regex:'conv_gemm2|conv_wgrad2' -c 6 -o gpurun_out/${w}_$tag -f python tools/aspp_bench.py 1 $w > gpurun_out/ncu_${w}_$tag.log 2>&1; echo "[ncu $w exit $?]" >> gpurun_out/ncu_${w}_$tag.log
      python tools/ncu_summary.py gpurun_out/${w}_$tag.ncu-rep > gpurun_out/${w}_${tag}_summary.txt 2>&1
      grep -E "tensor_cycles_active|time_duration|kernel:|dram__bytes_read.sum |lts__throughput|sm__cycles_elapsed.max.per_second|mem_tensor" gpurun_out/${w}_${tag}_summary.txt | head -60
    done
    ;;
  altdir)     # alternating tile directions (knob 12): bit-exactness test, kernel / network tests with the knob on, bench A/B on one box
    timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider -k "alternating" > gpurun_out/pytest_${tag}_alt.log 2>&1; tail -3 gpurun_out/pytest_${tag}_alt.log | cut -c1-200
    B200SEG_ALT_DIR=1 timeout -s KILL 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_graph.py tests/test_gpu_nets.py tests/test_gpu_fullsize.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_${tag}_alt1.log 2>&1; tail -3 gpurun_out/pytest_${tag}_alt1.log | cut -c1-200
    for v in 1 0 1 0; do
      B200SEG_ALT_DIR=$v B200SEG_SKIP_EXTRAS=1 B200SEG_SKIP_CPU_BASELINE=1 B200SEG_SHAPE_PROFILE=gpurun_out/shape_profile_${tag}_alt$v.txt bench_line ${tag}_alt$v --steps 10 --warmup 3 --no-second-precision --no-tf32-peak
    done
    paste -d'|' <(head -30 gpurun_out/shape_profile_${tag}_alt1.txt | cut -c1-100) <(head -30 gpurun_out/shape_profile_${tag}_alt0.txt | cut -c60-100)
    ;;
  wgbal)      # load-balanced weight-gradient plan + ring depth A/B (micro), wgrad / network tests with the new plan, bench
    timeout -s KILL 600 python tools/aspp_bench.py 5 wg > gpurun_out/micro_${tag}_wg.log 2>&1; cat gpurun_out/micro_${tag}_wg.log | cut -c1-110
    timeout -s KILL 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py tests/test_gpu_nets.py tests/test_gpu_graph.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_${tag}.log 2>&1; tail -3 gpurun_out/pytest_${tag}.log | cut -c1-200
    for st in 6 7; do
      B200SEG_WGRAD_STAGES=$st B200SEG_SKIP_EXTRAS=1 B200SEG_SKIP_CPU_BASELINE=1 B200SEG_SHAPE_PROFILE=gpurun_out/shape_profile_${tag}_st$st.txt bench_line ${tag}_st$st --steps 10 --warmup 3 --no-second-precision --no-tf32-peak
    done
    grep wgrad gpurun_out/shape_profile_${tag}_st6.txt | head -12
    ;;
  wgscan)     # (pixel splits, tap rotation) scan of the ASPP weight gradients + DRAM bytes of the d36 launch under the old / new plan
    timeout -s KILL 600 python tools/aspp_bench.py 4 wgscan > gpurun_out/micro_${tag}_wgscan.log 2>&1; cat gpurun_out/micro_${tag}_wgscan.log | cut -c1-110
    timeout -s KILL 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:conv_wgrad2 --csv --log-file gpurun_out/wgscan_dram_$tag.csv python tools/aspp_bench.py 1 wgscan > /dev/null 2>&1
    python - <<PYEOF
import csv
rows=[r for r in csv.reader(open('gpurun_out/wgscan_dram_$tag.csv')) if len(r)>10]
h=rows[0]; im=h.index('Metric Name'); iv=h.index('Metric Value'); iid=h.index('ID')
cur={}
for r in rows[1:]:
    cur.setdefault(r[iid],{})[r[im]]=r[iv]
for k,v in list(cur.items())[:120:2]:
    print(k, v)
PYEOF
    ;;
  libab)      # A/B of two builds of the library on one box (cutmix_semisup_seg_b200/libb200seg_prev.so = the previous commit): tests on the new one, bench prev / new / prev / new
    timeout -s KILL 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py tests/test_gpu_nets.py tests/test_gpu_graph.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_${tag}.log 2>&1; tail -3 gpurun_out/pytest_${tag}.log | cut -c1-200
    for v in prev new prev new; do
      lib=; [ $v = prev ] && lib=$PWD/cutmix_semisup_seg_b200/libb200seg_prev.so
      B200SEG_LIB=$lib B200SEG_SKIP_EXTRAS=1 B200SEG_SKIP_CPU_BASELINE=1 B200SEG_SHAPE_PROFILE=gpurun_out/shape_profile_${tag}_$v.txt bench_line ${tag}_$v --steps 10 --warmup 3 --no-second-precision --no-tf32-peak
    done
    paste -d'|' <(head -34 gpurun_out/shape_profile_${tag}_new.txt | cut -c1-100) <(head -34 gpurun_out/shape_profile_${tag}_prev.txt | cut -c60-100)
    ;;
  regate)     # BatchNorm backward with the gate recomputed from x: bit-exactness + whole netops / network files, bench A/B (B200SEG_BN_REGATE)
    timeout -s KILL 900 python -m pytest tests/test_gpu_netops.py tests/test_gpu_kernels.py tests/test_gpu_fullsize.py tests/test_gpu_nets.py tests/test_gpu_graph.py tests/test_zzy_gpu_denseunet.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_${tag}.log 2>&1; tail -3 gpurun_out/pytest_${tag}.log | cut -c1-200
    for v in 0 1 0 1; do
      B200SEG_BN_REGATE=$v B200SEG_SKIP_EXTRAS=1 B200SEG_SKIP_CPU_BASELINE=1 B200SEG_SHAPE_PROFILE=gpurun_out/shape_profile_${tag}_regate$v.txt bench_line ${tag}_regate$v --steps 10 --warmup 3 --no-second-precision --no-tf32-peak
    done
    grep -h "bn_bwd" gpurun_out/shape_profile_${tag}_regate0.txt gpurun_out/shape_profile_${tag}_regate1.txt
    ;;
  final)      # what the driver runs at round end (whole suite, smoke, default bench) + launch list of one eager step + the other bench lines
    run_tests $tag tests
    grep "fullsize cfg3 3xtf32\|fullsize cfg2 3xtf32 iteration 0" gpurun_out/parity_$tag.txt | cut -c1-230
    timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/smoke_$tag.log 2>&1; tail -2 gpurun_out/smoke_$tag.log | cut -c1-200
    B200SEG_SHAPE_PROFILE=gpurun_out/shape_profile_$tag.txt bench_line $tag
    timeout -s KILL 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/launches_$tag.csv python tools/ncu_step.py > gpurun_out/ncu_step_$tag.log 2>&1; echo "[ncu step exit $?]" >> gpurun_out/ncu_step_$tag.log
    python tools/launch_list_summary.py gpurun_out/launches_$tag.csv > gpurun_out/launch_list_summary_$tag.txt 2>&1; head -14 gpurun_out/launch_list_summary_$tag.txt
    gzip -f gpurun_out/launches_$tag.csv
    for spec in "aug:--loss aug" "vat:--loss vat" "ict:--loss ict" "config4:--arch denseunet --loss aug" "cfg2:--arch v2"; do
      n=${spec%%:*}; a=${spec#*:}
      B200SEG_SKIP_EXTRAS=1 B200SEG_SKIP_CPU_BASELINE=1 bench_line ${tag}_$n --steps 10 --warmup 3 --no-second-precision --no-tf32-peak $a
    done
    ;;
  ddp2b)      # 2 GPUs on HEAD: correctness of the bucketed / overlapped all-reduce with the batched head, one bench line
    timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py > gpurun_out/ddp_check_$tag.log 2>&1
    grep -E "buckets=|overlapped|DDP_CHECK|Error|error" gpurun_out/ddp_check_$tag.log | head -20
    B200SEG_SKIP_CPU_BASELINE=1 timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_${tag}_2gpu.log 2>&1
    grep '^{' gpurun_out/bench_${tag}_2gpu.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('2 GPUs', d['value'], 'img/s', d['ms_per_step'], 'ms e2e', d['e2e']['value'], d['clocks'])" || tail -5 gpurun_out/bench_${tag}_2gpu.log
    ;;
  occ2)       # after the cheap tile decoding: ncu of the layer3 3x3 / ASPP launches (occupancy = sm__mem_tensor_cycles_active) + clock64 pipeline trace
    for w in l3x3 aspp32; do
      timeout -s KILL 600 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.avg,sm__cycles_elapsed.max,sm__cycles_elapsed.avg.per_second,sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'conv_gemm2|conv_wgrad2' -c 6 --csv --log-file gpurun_out/occ_${w}_$tag.csv python tools/aspp_bench.py 1 $w > gpurun_out/ncu_occ_${w}_$tag.log 2>&1
      python - <<PYEOF
import csv
rows=[r for r in csv.reader(open('gpurun_out/occ_${w}_$tag.csv')) if len(r)>10]
h=rows[0]; im=h.index('Metric Name'); iv=h.index('Metric Value'); iid=h.index('ID'); ik=h.index('Kernel Name')
cur={}
for r in rows[1:]:
    cur.setdefault((r[iid], r[ik][:40]),{})[r[im]]=r[iv]
for k,v in cur.items(): print('$w', k, v)
PYEOF
    done
    timeout -s KILL 300 python tools/aspp_bench.py 3 trace3 > gpurun_out/trace3_$tag.log 2>&1; cat gpurun_out/trace3_$tag.log | cut -c1-250
    timeout -s KILL 300 python tools/aspp_bench.py 5 all > gpurun_out/micro_${tag}_all.log 2>&1; cat gpurun_out/micro_${tag}_all.log | cut -c1-100
    ;;
  epi2slot)   # one-operand TMA epilogue flavours with two requests in flight: bit-exactness + timing per flavour, kernel tests, bench prev / new
    timeout -s KILL 300 python tools/aspp_bench.py 3 tma > gpurun_out/micro_${tag}_tma.log 2>&1; echo "[tma exit $?]" >> gpurun_out/micro_${tag}_tma.log; cut -c1-200 gpurun_out/micro_${tag}_tma.log
    B200SEG_LIB=$PWD/cutmix_semisup_seg_b200/libb200seg_prev.so timeout -s KILL 300 python tools/aspp_bench.py 3 tma > gpurun_out/micro_${tag}_tma_prev.log 2>&1; grep "residual\|gate only" gpurun_out/micro_${tag}_tma_prev.log | cut -c1-120
    timeout -s KILL 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py tests/test_gpu_nets.py tests/test_gpu_graph.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_${tag}.log 2>&1; tail -3 gpurun_out/pytest_${tag}.log | cut -c1-200
    for v in prev new prev new; do
      lib=; [ $v = prev ] && lib=$PWD/cutmix_semisup_seg_b200/libb200seg_prev.so
      B200SEG_LIB=$lib B200SEG_SKIP_EXTRAS=1 B200SEG_SKIP_CPU_BASELINE=1 B200SEG_SHAPE_PROFILE=gpurun_out/shape_profile_${tag}_$v.txt bench_line ${tag}_$v --steps 10 --warmup 3 --no-second-precision --no-tf32-peak
    done
    paste -d'|' <(head -24 gpurun_out/shape_profile_${tag}_new.txt | cut -c1-100) <(head -24 gpurun_out/shape_profile_${tag}_prev.txt | cut -c60-100)
    ;;
  widepf)     # wide L2 prefetch of the A stream (knob 18 / B200SEG_WIDE_PF): micro A/B, bench A/B
    timeout -s KILL 300 python tools/aspp_bench.py 5 widepf > gpurun_out/micro_${tag}_widepf.log 2>&1; echo "[exit $?]" >> gpurun_out/micro_${tag}_widepf.log; cut -c1-120 gpurun_out/micro_${tag}_widepf.log
    for v in 0 1 0 1; do
      B200SEG_WIDE_PF=$v B200SEG_SKIP_EXTRAS=1 B200SEG_SKIP_CPU_BASELINE=1 B200SEG_SHAPE_PROFILE=gpurun_out/shape_profile_${tag}_pf$v.txt bench_line ${tag}_pf$v --steps 10 --warmup 3 --no-second-precision --no-tf32-peak
    done
    paste -d'|' <(head -16 gpurun_out/shape_profile_${tag}_pf1.txt | cut -c1-100) <(head -16 gpurun_out/shape_profile_${tag}_pf0.txt | cut -c60-100)
    ;;
  final2)     # whole suite + smoke + default bench on HEAD (no profiler passes)
    run_tests $tag tests
    grep "fullsize cfg3 3xtf32\|fullsize cfg2 3xtf32 iteration 0" gpurun_out/parity_$tag.txt | cut -c1-230
    timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/smoke_$tag.log 2>&1; tail -2 gpurun_out/smoke_$tag.log | cut -c1-200
    B200SEG_SHAPE_PROFILE=gpurun_out/shape_profile_$tag.txt bench_line $tag
    ;;
  im2col)     # shared-memory stem im2col: bit-exactness tests, bench A/B (B200SEG_IM2COL_GLOBAL=1 = the previous kernel)
    timeout -s KILL 600 python -m pytest tests/test_gpu_netops.py tests/test_gpu_nets.py tests/test_zz_gpu_vat.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_${tag}.log 2>&1; tail -3 gpurun_out/pytest_${tag}.log | cut -c1-200
    for v in 1 0 1 0; do
      gl=; [ $v = 1 ] && gl=1
      B200SEG_IM2COL_GLOBAL=$gl B200SEG_SKIP_EXTRAS=1 B200SEG_SKIP_CPU_BASELINE=1 B200SEG_SHAPE_PROFILE=gpurun_out/shape_profile_${tag}_global$v.txt bench_line ${tag}_global$v --steps 10 --warmup 3 --no-second-precision --no-tf32-peak
    done
    grep -h im2col gpurun_out/shape_profile_${tag}_global1.txt gpurun_out/shape_profile_${tag}_global0.txt
    ;;
  netops2)    # row-based resize / im2col kernels: bit-exactness vs the flat forms, netops + network tests, event timings, bench
    timeout -s KILL 600 python -m pytest tests/test_gpu_netops.py tests/test_gpu_nets.py tests/test_gpu_graph.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_${tag}.log 2>&1; tail -3 gpurun_out/pytest_${tag}.log | cut -c1-200
    B200SEG_SKIP_EXTRAS=1 B200SEG_SKIP_CPU_BASELINE=1 B200SEG_SHAPE_PROFILE=gpurun_out/shape_profile_${tag}.txt bench_line ${tag} --steps 10 --warmup 3 --no-second-precision --no-tf32-peak
    grep "bilinear\|im2col\|maxpool\|gap\|bn_" gpurun_out/shape_profile_${tag}.txt
    ;;
  micro)      timeout -s KILL 600 python tools/aspp_bench.py 3 ${3:-all} > gpurun_out/micro_$tag.log 2>&1; cat gpurun_out/micro_$tag.log | cut -c1-120 ;;
  bench)      shift 2; B200SEG_SHAPE_PROFILE=gpurun_out/shape_profile_$tag.txt bench_line $tag "$@" ;;
  *) echo "unknown stage $stage"; exit 2 ;;
esac

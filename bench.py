"""Benchmark of the CutMix mean-teacher training iteration (BASELINE.json metric: images/sec at 512x512,
bs=16 per GPU, DeepLab v3+ ResNet-101, 19 classes, synthetic Cityscapes-shaped data).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--arch v3plus|v2|denseunet] [--loss ...]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one full iteration of train_seg_semisup_mask_mt.py:287-476: supervised forward/backward on N
labelled images, two teacher forwards + one student forward/backward on 2N unlabelled images (CutMix),
fused losses, Adam step, EMA step, (N>1) one gradient all-reduce.  images/sec := N * world / iteration time.

Prints ONE JSON line (rank 0).  `value` is timed with inputs resident in HBM; `e2e` repeats the measurement
through the public API with pinned host batches copied H2D and the loss read back D2H every step.
`--impl reference` times the CPU port of the reference path (oracle/) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = {
    'v3plus': dict(arch='resnet101_deeplabv3plus_imagenet', classes=19, h=512, w=512, batch=16, lr=1e-5,
                   workload='DeepLab v3+ ResNet-101, synthetic Cityscapes 512x512, 19 classes, CutMix mean teacher, bs 16/GPU'),
    'v2': dict(arch='resnet101_deeplab_imagenet', classes=21, h=321, w=321, batch=16, lr=3e-5,
               workload='DeepLab v2 ResNet-101, synthetic Pascal-Aug 321x321, 21 classes, CutMix mean teacher, bs 16/GPU'),
    # BASELINE config 4 (run it with --loss aug: train_seg_semisup_aug_mt.py on the DenseNet-161 U-Net)
    'denseunet': dict(arch='densenet161unet', classes=2, h=224, w=224, batch=16, lr=1e-5,
                      workload='DenseNet-161 U-Net, synthetic ISIC2017 224x224, 2 classes, CutMix mean teacher, bs 16/GPU'),
}
# forward conv FLOPs per image (2*MAC, dense), stem FLOPs: SURVEY.md §8 / BASELINE.md §2
FLOPS = {'v3plus': (520.28e9, 1.233e9), 'v2': (147.67e9, 0.488e9),
         'denseunet': (37.15e9, 0.354e9)}      # counted from the graph's convolution launches (K un-padded), 224x224


def committed_traffic(kernel):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, mean over the launches of one eager step) of
    `kernel` from the newest committed ncu launch-list summary under profiles/ (tools/ncu_step.py + launch_list_summary.py).
    ncu cannot run inside the timed bench, so this is the committed measurement of the same step, not a live one."""
    import glob
    import re

    def version(path):          # profiles/rNN_vMM_...: newest round, then newest capture
        m = re.search(r'r(\d+)_v(\d+)_', os.path.basename(path))
        return (int(m.group(1)), int(m.group(2))) if m else (-1, -1)
    files = sorted(glob.glob(os.path.join(ROOT, 'profiles', '*_launch_list_summary.txt')), key=version)
    for path in reversed(files):
        try:
            for line in open(path):
                f = line.split()
                if len(f) >= 7 and f[0] == kernel:
                    return int((float(f[4]) + float(f[5])) * 1e6), os.path.relpath(path, ROOT)
        except (OSError, ValueError):
            continue
    return None, None


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d['hbm_gbs'], bf16_burst=d['bf16_tflops'], bf16_sustained=d['bf16_tflops_sustained'],
                    source='measured')
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source='fallback')


class ClockSampler(object):
    """nvidia-smi clocks/throttle sampler running during the timed region."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q, '--format=csv,noheader,nounits',
                                          '-lms', '200'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            try:
                pw.append(float(f[2]))
            except ValueError:
                pass
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm), 'power_w': float(np.median(pw)) if pw else None}


# ------------------------------------------------------------------------------------------------ B200 arm
def build_trainer(cfg, device, dist_on, use_graph=True, precision='tf32'):
    from architectures import network_architectures
    import mask_gen
    import optim_weight_ema
    from cutmix_semisup_seg_b200 import step as step_mod, synthetic
    torch.manual_seed(0)
    Net = network_architectures.seg.get(cfg['arch'])
    import inspect
    kw = dict(pretrained=False) if 'pretrained' in inspect.signature(Net).parameters else {}
    student = Net(cfg['classes'], **kw)
    synthetic.condition_classifier(student, 40.0)
    student = student.to(device)
    teacher = Net(cfg['classes'], **kw).to(device)
    student.b2_precision = teacher.b2_precision = precision        # 'tf32' (one tensor-core pass) | '3xtf32' (parity mode)
    for p in teacher.parameters():
        p.requires_grad = False
    optim = step_mod.make_optimizer(student, 'adam', cfg['lr'], capturable=use_graph,
                                    fused_kernel=os.environ.get('B200SEG_FUSED_OPT', '1') != '0')
    ema = optim_weight_ema.EMAWeightOptimizer(teacher, student, 0.99)
    student.train(); teacher.train()
    student.freeze_batchnorm(); teacher.freeze_batchnorm()         # every reference recipe uses --freeze_bn
    mg = mask_gen.BoxMaskGenerator(prop_range=0.5, n_boxes=1, random_aspect_ratio=True, prop_by_area=True,
                                   within_bounds=True, invert=True)
    trainer = step_mod.MeanTeacherStep(student, teacher, optim, ema, mg, cons_loss_fn='var', cons_weight=1.0,
                                       conf_thresh=0.97, conf_per_pixel=False, rampup=-1, mask_mix=True,
                                       dist_group=True if dist_on else None, use_cuda_graph=use_graph,
                                       batch_trunk=os.environ.get('B200SEG_BATCH_TRUNK', '1') != '0')
    return trainer, mg


def timed_conv_profile(trainer, sup, unsup):
    """One instrumented iteration: CUDA-event duration of every tensor-core conv launch (on the launching
    stream) and its algorithmic FLOPs -> aggregate achieved TFLOP/s per kernel."""
    be = trainer.be
    be.start_profile()
    trainer.step(sup, [unsup], eager=True)
    torch.cuda.synchronize()
    return be.stop_profile()


def in_graph_profile(trainer, sup, unsup, iters=3):
    """Device timeline of graph-replayed iterations through CUPTI (torch.profiler; outside the timed regions): per-kernel time
    INSIDE the replay, where kernels run back to back at the sustained (power-capped) clock, and the device idle time between
    them.  Returns {'kernels': {short name: [launches, ms]} per iteration, 'span_ms', 'idle_ms'} of the middle iteration."""
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(iters):
            trainer.step(sup, [unsup])
        torch.cuda.synchronize()
    evs = sorted((e.time_range.start, e.time_range.end, e.name) for e in prof.events()
                 if e.device_type == torch.autograd.DeviceType.CUDA)
    per = len(evs) // iters
    if per == 0:
        return None
    it = evs[per:2 * per] if iters >= 3 else evs[:per]
    kernels, busy_end, idle = {}, it[0][0], 0.0
    for s, e, name in it:
        short = name.replace('void ', '').replace('(anonymous namespace)::', '').split('(')[0].split('<')[0]
        k = kernels.setdefault(short, [0, 0.0]); k[0] += 1; k[1] += (e - s) / 1e3
        if s > busy_end:
            idle += (s - busy_end) / 1e3
        busy_end = max(busy_end, e)
    return {'kernels': kernels, 'span_ms': (busy_end - it[0][0]) / 1e3, 'idle_ms': idle, 'activities': per}


def _release():
    """After the caller dropped its trainer (CUDA graphs, tens of GB of graph-private activations): return the memory before
    the next leg builds its own."""
    import gc
    gc.collect()
    torch.cuda.empty_cache()


def timed_steps(trainer, sup_dev, uns_dev, steps, warmup, sync_all):
    """W untimed + K timed iterations with inputs resident in HBM; CUDA events on the launching stream."""
    pool = len(sup_dev)
    for i in range(warmup):
        out = trainer.step(sup_dev[i % pool], [uns_dev[i % pool]])
    sync_all()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = trainer.be.launches
    t_start = time.time()
    ev0.record()
    for i in range(steps):
        out = trainer.step(sup_dev[i % pool], [uns_dev[i % pool]])
    ev1.record()
    t_enq = time.time()
    sync_all()
    return ev0.elapsed_time(ev1), trainer.be.launches - launches0, out, (t_enq - t_start)


def parity_block(device):
    """Losses of the full-size cfg3 parity iterations (tests/fullsize_recipe.py: DeepLab v3+, N = 16, 512 x 512, seeded weights /
    batches / dropout masks) in the precision this bench times (single-pass TF32) and in the parity mode (3xTF32), next to the
    values the UNMODIFIED reference modules produced on the CPU (tests/golden/fullsize_cfg3.npz, committed fixture written by
    oracle/gen_golden_fullsize.py).  Eager iterations outside every timed region."""
    import warnings
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import fullsize_recipe as R
    from architectures import network_architectures
    import mask_gen
    import optim_weight_ema
    from cutmix_semisup_seg_b200 import step as step_mod, synthetic
    name = 'cfg3'
    path = os.path.join(ROOT, 'tests', 'golden', 'fullsize_%s.npz' % name)
    if not os.path.exists(path):
        return {'unavailable': 'tests/golden/fullsize_%s.npz missing' % name}
    gold = np.load(path)
    cfg = R.CONFIGS[name]
    ref = {k: [float(v) for v in np.atleast_1d(gold[k])] for k in ('sup_loss', 'cons_loss', 'conf_rate')}
    out = {'config': 'DeepLab v3+ ResNet-101, N = {n}, {h}x{w}, {classes} classes, conf_thresh {conf_thresh}, Adam lr {lr}, '
                     '{iters} consecutive iterations, injected dropout masks'.format(**cfg),
           'reference': dict(ref, source='tests/golden/fullsize_cfg3.npz: unmodified reference modules, torch CPU fp32')}
    runs = {}
    for precision in ('tf32', '3xtf32'):
        Net = network_architectures.seg.get(cfg['kind'])
        student = Net(cfg['classes'], pretrained=False)
        sd = synthetic.synth_state_dict(student.state_dict(), seed=cfg['seed'], logit_gain=cfg['gain'],
                                        final_keys=R.final_keys(student.state_dict(), cfg))
        student.load_state_dict(sd)
        teacher = Net(cfg['classes'], pretrained=False)
        student.to(device); teacher.to(device)
        student.b2_precision = teacher.b2_precision = precision
        for p in teacher.parameters():
            p.requires_grad = False
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            optim = step_mod.make_optimizer(student, 'adam', cfg['lr'], fused_kernel=True)
        ema = optim_weight_ema.EMAWeightOptimizer(teacher, student, 0.99)
        student.train(); teacher.train(); student.freeze_batchnorm(); teacher.freeze_batchnorm()
        mg = mask_gen.BoxMaskGenerator(0.5, invert=True)
        tr = step_mod.MeanTeacherStep(student, teacher, optim, ema, mg, conf_thresh=cfg['conf_thresh'])
        vals = {'sup_loss': [], 'cons_loss': [], 'conf_rate': []}
        for it in range(cfg['iters']):
            (sx, sy), uns = R.batches(cfg, mg, compact_masks=True, it=it)
            dm = R.dropout_masks(cfg, it)
            for net, keys in ((student, ('sup', 'stu')), (teacher, ('tea0', 'tea1'))):
                drop = [m for m in net.modules() if type(m).__name__ == 'B2Dropout'][0]
                drop.inject([dm[k] for k in keys])
            o = tr.step((sx.to(device), sy.to(device)), [{k: v.to(device) for k, v in uns.items()}])
            for k in vals:
                vals[k].append(float(o[k]))
        runs[precision] = vals
        tr = student = teacher = optim = ema = None
        _release()

    def dev(a, b):
        return {'sup_loss_rel': max(abs(x - y) / abs(y) for x, y in zip(a['sup_loss'], b['sup_loss'])),
                'cons_loss_rel': max(abs(x - y) / abs(y) for x, y in zip(a['cons_loss'], b['cons_loss'])),
                'conf_rate_abs': max(abs(x - y) for x, y in zip(a['conf_rate'], b['conf_rate']))}
    for precision in runs:
        out[precision] = dict(runs[precision], max_dev_vs_reference=dev(runs[precision], ref))
    out['tf32_vs_3xtf32'] = dev(runs['tf32'], runs['3xtf32'])
    return out


def incumbent_cudnn(args, cfg, device, n, steps):
    """The honest incumbent (BASELINE.md section 3 "Also reported"): the reference's own path -- ATen / cuDNN convolutions under
    torch.autograd, torch.optim.Adam, the EMA loop -- on THIS GPU: oracle/ref_step.py (the line-for-line restatement of
    train_seg_semisup_mask_mt.py:287-476 on functional torch networks, pinned to the reference by tests/golden/) with its
    state on the device.  Reported next to `value`; none of this repository's kernels run in it.  Three settings: the stock
    one (cudnn.allow_tf32 = True, PyTorch's default, cudnn.benchmark = False as the reference never sets it), full fp32
    (allow_tf32 = False) and stock + cudnn.benchmark = True (cuDNN's autotuned best)."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ref_step
    import inspect
    from architectures import network_architectures
    import mask_gen
    from cutmix_semisup_seg_b200 import synthetic
    ctor = network_architectures.seg.get(cfg['arch'])
    kw = dict(pretrained=False) if 'pretrained' in inspect.signature(ctor).parameters else {}
    net = ctor(cfg['classes'], **kw)
    final = [k for k in net.state_dict() if ('layer5' in k or 'classifier.classifier.6' in k or 'final_clf' in k) and k.endswith('weight')]
    sd = synthetic.synth_state_dict(net.state_dict(), seed=0, logit_gain=12.0, final_keys=final)
    del net
    arch = {'v3plus': 'deeplab3plus', 'v2': 'deeplab2', 'denseunet': 'denseunet'}[args.arch]
    mg = mask_gen.BoxMaskGenerator(prop_range=0.5, n_boxes=1, random_aspect_ratio=True, prop_by_area=True,
                                   within_bounds=True, invert=True)
    h, w = cfg['h'], cfg['w']
    sup = synthetic.make_sup_batch(n, h, w, cfg['classes'], 100, device=device)
    uns = synthetic.make_unsup_batch(n, h, w, 200, mg, compact_masks=False, device=device)
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    res = {'what': 'reference path (ATen/cuDNN + autograd + torch.optim.Adam(foreach=False) + EMA loop, NCHW fp32 tensors) on the same '
                   'B200, batch {} at {}x{}'.format(n, h, w), 'unit': 'images/s', 'steps': steps}
    try:
        for tag, tf32, bench in (('tf32_stock', True, False), ('fp32', False, False), ('tf32_cudnn_benchmark', True, True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.benchmark = bench
            tr = ref_step.OracleMeanTeacher(arch, {k: v.to(device) for k, v in sd.items()}, cfg['lr'])

            def drop():
                if arch != 'deeplab3plus':
                    return None
                mk = lambda: [(torch.rand((n, 256, -(-h // 8), -(-w // 8)), device=device) > 0.5).float()]   # noqa: E731
                return {'sup': mk(), 'tea0': mk(), 'tea1': mk(), 'stu': mk()}
            for _ in range(2):
                tr.step(sup[0], sup[1], uns, drop=drop())
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                last = tr.step(sup[0], sup[1], uns, drop=drop())
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            res[tag] = {'value': round(n / (ms / 1e3), 3), 'ms_per_step': round(ms, 2), 'last_losses': [round(float(v), 6) for v in last]}
            tr = None
            _release()
    except Exception as e:      # keep the bench line if torch's own path fails (e.g. out of memory)
        res['failed'] = repr(e)[:300]
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
    return res


def run_b200(args):
    rank = int(os.environ.get('RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py (b200 arm) needs a CUDA device: there is no CPU fallback')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    dist_on = world > 1
    if dist_on:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        import datetime
        # fail fast on a rank mismatch instead of NCCL's 10 minute default (a hung box burns the GPU budget)
        dist.init_process_group('nccl', device_id=device, timeout=datetime.timedelta(seconds=180))
    from cutmix_semisup_seg_b200 import synthetic
    cfg = CFG[args.arch]
    n, h, w = args.batch or cfg['batch'], cfg['h'], cfg['w']
    trainer, mg = build_trainer(cfg, device, dist_on, use_graph=not args.eager, precision=args.precision)
    be = trainer.be
    # a small pool of distinct batches (per-iteration working set >> 126 MB L2: activations alone are ~20 GB)
    pool = 3
    sup_host = [synthetic.make_sup_batch(n, h, w, cfg['classes'], 100 + rank * 10 + i, pin=True) for i in range(pool)]
    if args.loss == 'ict':       # train_seg_semisup_ict.py iteration: per-sample Beta mix factors
        uns_host = [synthetic.make_ict_batch(n, h, w, 200 + rank * 10 + i, 0.1, pin=True) for i in range(pool)]
    elif args.loss == 'aug':     # train_seg_semisup_aug_mt.py iteration: two geometrically different views + affine maps
        uns_host = [synthetic.make_aug_batch(n, h, w, 200 + rank * 10 + i, rot_mag=10.0, max_scale=1.2, offset_range=16.0,
                                             pin=True) for i in range(pool)]
    elif args.loss == 'vat':     # train_seg_semisup_vat_mt.py iteration: adversarial perturbation from an input-gradient pass
        uns_host = [synthetic.make_vat_batch(n, h, w, 200 + rank * 10 + i, pin=True) for i in range(pool)]
    else:
        uns_host = [synthetic.make_unsup_batch(n, h, w, 200 + rank * 10 + i, mg, pin=True) for i in range(pool)]
    sup_dev = [(a.to(device), b.to(device)) for a, b in sup_host]
    uns_dev = [{k: v.to(device) for k, v in d.items()} for d in uns_host]

    def sync_all():
        torch.cuda.synchronize()
        if dist_on:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(args.warmup):
        out = trainer.step(sup_dev[i % pool], [uns_dev[i % pool]])
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches, out, enq_s = timed_steps(trainer, sup_dev, uns_dev, args.steps, 0, sync_all)
    clocks = sampler.stop() if rank == 0 else None
    last = {k: float(v) for k, v in out.items() if v is not None}

    # ---- end to end: pinned host batches -> H2D every step, loss scalars -> D2H every step
    h2d = sum(t.numel() * t.element_size() for t in sup_host[0]) + \
        sum(t.numel() * t.element_size() for t in {id(v): v for v in uns_host[0].values()}.values())
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    d2h = 0
    for i in range(args.steps):
        # pinned host batches go through the public step() call: every step's H2D copy is inside the timed region -- the
        # first one directly into the graph's static buffers, the later ones prefetched on a side stream during the
        # previous step (`prefetch=`) and moved device-to-device at the start of their step
        nxt = (sup_host[(i + 1) % pool], [uns_host[(i + 1) % pool]]) if i + 1 < args.steps else None
        o = trainer.step(sup_host[i % pool], [uns_host[i % pool]], prefetch=nxt)    # next batch's H2D overlaps this step
        vals = torch.stack([o['sup_loss'], o['cons_loss'], o['conf_rate']]).cpu()    # D2H read of the step's result
        d2h = vals.numel() * 4
    e1.record()
    sync_all()
    ms_e2e = e0.elapsed_time(e1)

    # ---- optional: the same end-to-end loop with uint8 batches crossing PCIe: images as HWC uint8, labels
    # and valid masks as uint8, standardised / widened on the device by csrc/input.cu (DeviceNormalizeToTensor)
    e2e_u8 = None
    if args.u8_inputs and args.loss == 'cutmix':
        from cutmix_semisup_seg_b200 import input_pipeline
        norm = input_pipeline.DeviceNormalizeToTensor([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])

        def to_u8_img(t):        # synthetic standardised images -> plausible uint8 pixels (HWC), pinned
            return (t * 58.0 + 116.0).clamp_(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous().pin_memory()
        u8_pool = []
        for i in range(pool):
            sx, sy = sup_host[i]
            ub = uns_host[i]
            u8_pool.append(dict(sup=dict(image_arr=to_u8_img(sx), labels_arr=sy[:, 0].to(torch.uint8).pin_memory()),
                                v0=dict(image_arr=to_u8_img(ub['ux0_tea']), mask_arr=(ub['um0'][:, 0] * 255).to(torch.uint8).pin_memory()),
                                v1=dict(image_arr=to_u8_img(ub['ux1_tea']), mask_arr=(ub['um1'][:, 0] * 255).to(torch.uint8).pin_memory()),
                                mask_params=ub['mask_params']))
        h2d_u8 = sum(t.numel() * t.element_size() for d in (u8_pool[0]['sup'], u8_pool[0]['v0'], u8_pool[0]['v1']) for t in d.values()) \
            + u8_pool[0]['mask_params'].numel() * u8_pool[0]['mask_params'].element_size()

        def u8_step(b):
            s_ = norm(b['sup']); a0 = norm(b['v0']); a1 = norm(b['v1'])
            uns = dict(ux0_tea=a0['image'], ux0_stu=a0['image'], um0=a0['mask'], ux1_tea=a1['image'], ux1_stu=a1['image'],
                       um1=a1['mask'], mask_params=b['mask_params'].to(device, non_blocking=True))
            return trainer.step((s_['image'], s_['labels']), [uns])
        for i in range(2):
            u8_step(u8_pool[i % pool])
        sync_all()
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        u0.record()
        for i in range(args.steps):
            o = u8_step(u8_pool[i % pool])
            torch.stack([o['sup_loss'], o['cons_loss'], o['conf_rate']]).cpu()
        u1.record()
        sync_all()
        e2e_u8 = {'value': round(n * world / (u0.elapsed_time(u1) / args.steps / 1e3), 3), 'unit': 'images/s',
                  'h2d_bytes_per_step': int(h2d_u8), 'note': 'uint8 batches, normalise-to-tensor on the device; no prefetch overlap'}

    # ---- per-kernel roofline (instrumented iteration outside the timed regions)
    # every rank runs it (the iteration contains the gradient all-reduce); only rank 0 reports
    prof = timed_conv_profile(trainer, sup_dev[0], uns_dev[0])
    in_graph = None
    if not args.eager and not os.environ.get('B200SEG_SKIP_IN_GRAPH'):
        try:
            for _ in range(2):
                trainer.step(sup_dev[0], [uns_dev[0]])
            in_graph = in_graph_profile(trainer, sup_dev[0], uns_dev[0])
        except Exception as e:          # CUPTI unavailable: the eager per-launch profile stands alone
            in_graph = {'failed': repr(e)[:200]}
    if rank != 0:
        prof = None
    shape_profile = getattr(be, 'last_shape_profile', None)
    optim_note = trainer.optim_note
    trunk_batched = trainer._can_batch_trunk([None])
    trainer = None
    _release()

    # ---- the other precision mode on the same workload (fewer steps: it only has to place the parity mode's cost)
    other = None
    other_mode = '3xtf32' if args.precision == 'tf32' else 'tf32'
    if not args.no_second_precision and world == 1:
        tr2, _ = build_trainer(cfg, device, dist_on, use_graph=not args.eager, precision=other_mode)
        k2 = max(3, min(args.steps, 5))
        ms2, _, out2, _ = timed_steps(tr2, sup_dev, uns_dev, k2, 3, sync_all)
        other = (ms2 / k2, k2, {k: float(v) for k, v in out2.items() if v is not None})
        tr2 = None
        _release()

    if dist_on:
        t = torch.tensor([ms, ms_e2e], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if dist_on:
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    ms_step = ms / args.steps
    value = n * world / (ms_step / 1e3)
    F, Fs = FLOPS[args.arch]
    # CutMix / ICT: 4 forward + 2 backward passes; augmentation consistency runs the teacher once (3 forward passes)
    # VAT: 5 forward passes (direction net twice, teacher, student twice), 2 full backward passes + 1 data-gradient-only pass
    flops_iter = ({'aug': 7, 'vat': 10}.get(args.loss, 8) * F - 2 * Fs) * n
    res = {
        'metric': 'images/sec', 'value': round(value, 3), 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': round(ms_step, 3), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': args.precision, 'data': 'synthetic',
        'config': {'workload': cfg['workload'].replace('CutMix', {'ict': 'ICT', 'aug': 'augmentation-consistency', 'vat': 'VAT'}.get(args.loss, 'CutMix')),
                   'global_batch': n * world, 'crop': [h, w], 'parallelism': 'dp%d' % world,
                   'precision': {'tf32': 'single-pass kind::tf32 tensor-core products, fp32 accumulate (PyTorch / cuDNN default for '
                                         'convolutions); parity of this mode: see `parity`',
                                 '3xtf32': 'three tensor-core passes over hi/lo operand splits (~fp32 products): the mode the tight '
                                           'parity tests run in'}[args.precision],
                   'l2': 'per-iteration working set (activations ~GBs) far exceeds the 126 MB L2; 3 distinct batches rotate',
                   'freeze_bn': True, 'optimizer': optim_note,
                   'trunk_batching': 'frozen-BN backbone once per network over 2 concatenated mini-batches'
                                     if trunk_batched else 'pass by pass', 'launch_mode': 'eager' if args.eager else 'cuda-graph replay (2 graphs/step)',
                   'host_enqueue_ms_per_step': round(enq_s * 1e3 / args.steps, 2),
                   'conv_tflops_per_s_whole_step': round(flops_iter / (ms_step / 1e3) / 1e12, 2)},
        'clocks': clocks,
        'e2e': {'value': round(n * world / (ms_e2e / args.steps / 1e3), 3), 'unit': 'images/s',
                'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h)},
        'gpu_launches': int(launches),
        **({'e2e_u8': e2e_u8} if e2e_u8 is not None else {}),
        'last_step': last,
    }
    modes = {args.precision: {'value': round(value, 3), 'ms_per_step': round(ms_step, 3), 'steps': args.steps}}
    if other:
        modes[other_mode] = {'value': round(n * world / (other[0] / 1e3), 3), 'ms_per_step': round(other[0], 3), 'steps': other[1],
                             'last_step': other[2]}
    res['precision_modes'] = modes
    if prof:
        dom = max(((k, v) for k, v in prof.items() if v['flops'] > 0), key=lambda kv: kv[1]['ms'])
        name, d = dom
        ach = d['flops'] / (d['ms'] / 1e3) / 1e12
        traffic, traffic_src = committed_traffic(name)
        tf32_peak = peaks.get('tf32_sustained')
        res['roofline'] = {'kernel': name, 'bound': 'tensor', 'achieved': round(ach, 2), 'peak': peaks['bf16_sustained'],
                           'unit': 'TFLOP/s', 'frac': round(ach / peaks['bf16_sustained'], 4), 'traffic': traffic,
                           'traffic_source': traffic_src,
                           'peak_source': peaks['source'] + ' cuBLAS bf16 sustained (MEASURED_PEAKS.json holds no tf32 figure; kind::tf32 '
                                          'issues at half the bf16 rate, so 0.5 is this kernel family\'s ceiling on this scale)',
                           'frac_of_half_bf16_rate': round(ach / (peaks['bf16_sustained'] / 2), 4),
                           'launches': d['n'], 'kernel_ms_per_step': round(d['ms'], 3),
                           'share_of_step': round(d['ms'] / ms_step, 4),
                           'profiled_step_kernel_ms_total': round(sum(v['ms'] for v in prof.values()), 3),
                           'per_kernel': {k: {'ms': round(v['ms'], 3), 'n': v['n'],
                                              'tflops': round(v['flops'] / max(v['ms'], 1e-9) / 1e9, 2)}
                                          for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms'])[:24]}}
        if in_graph and 'kernels' in in_graph:
            # the same kernel INSIDE the graph replay (CUPTI durations): back to back at the sustained clock, i.e. what the timed
            # region actually contains; `achieved` above comes from event-timed eager launches that run cooler
            ig = in_graph['kernels'].get(name)
            if ig:
                res['roofline']['in_graph'] = {
                    'kernel_ms_per_step': round(ig[1], 3), 'launches': ig[0], 'achieved': round(d['flops'] / (ig[1] / 1e3) / 1e12, 2),
                    'frac': round(d['flops'] / (ig[1] / 1e3) / 1e12 / peaks['bf16_sustained'], 4),
                    'share_of_step': round(ig[1] / in_graph['span_ms'], 4), 'iteration_span_ms': round(in_graph['span_ms'], 3),
                    'device_idle_ms': round(in_graph['idle_ms'], 3), 'device_activities': in_graph['activities'],
                    'how': 'CUPTI kernel durations of one graph-replayed iteration (torch.profiler), outside the timed region'}
        elif in_graph:
            res['roofline']['in_graph'] = in_graph
        if args.measure_tf32_peak:
            # MEASURED_PEAKS.json carries bf16 only; the tf32 library rate on THIS box at the clock it settles to, measured
            # here (outside the timed regions), is the like-for-like denominator of a kind::tf32 kernel
            m = measure_tf32_matmul(device)
            res['roofline']['tf32_matmul_measured'] = m
            res['roofline']['frac_of_measured_tf32_matmul'] = round(ach / m['tflops_sustained'], 4)
    if prof and os.environ.get('B200SEG_SHAPE_PROFILE') and shape_profile:
        top = sorted(shape_profile.items(), key=lambda kv: -kv[1]['ms'])[:70]
        with open(os.environ['B200SEG_SHAPE_PROFILE'], 'w') as f:
            for k, v in top:
                f.write('{:<70s} {:8.3f} ms  n={:4d}  {:7.1f} TFLOP/s\n'.format(k, v['ms'], v['n'], v['flops'] / max(v['ms'], 1e-9) / 1e9))
    extras = world == 1 and not os.environ.get('B200SEG_SKIP_EXTRAS')
    if extras and args.arch == 'v3plus' and args.loss == 'cutmix':
        try:
            res['parity'] = parity_block(device)
        except Exception as e:
            res['parity'] = {'failed': repr(e)[:300]}
    if extras and args.loss == 'cutmix':
        res['incumbent_cudnn'] = incumbent_cudnn(args, cfg, device, n, max(2, min(args.steps, 5)))
    if world > 1:       # the host baseline is a property of the box, reported by the N = 1 run only
        res['cpu_baseline'] = {'value': None, 'unit': 'images/s', 'cores': 0, 'kind': 'port', 'sample': 'reported at N=1 only'}
    elif os.environ.get('B200SEG_SKIP_CPU_BASELINE'):
        res['cpu_baseline'] = {'value': None, 'unit': 'images/s', 'cores': 0, 'kind': 'port', 'sample': 'skipped (B200SEG_SKIP_CPU_BASELINE)'}
    else:
        res['cpu_baseline'] = cpu_baseline(args, sample_batch=2)
    print(json.dumps(res))
    if dist_on:
        dist.destroy_process_group()


def measure_tf32_matmul(device):
    """cuBLAS fp32 matmul with TF32 tensor cores allowed, 8192^3, sustained for ~2 s: the library's tf32 rate on this GPU at the
    clock it settles to (MEASURED_PEAKS.json has bf16 only)."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn((8192, 8192), device=device); b = torch.randn((8192, 8192), device=device)
        for _ in range(5):
            torch.matmul(a, b)
        torch.cuda.synchronize()
        best, n_it, t_end = 0.0, 0, time.time() + 2.0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        while time.time() < t_end:
            for _ in range(20):
                torch.matmul(a, b)
            n_it += 20
            torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        sustained = 2 * 8192 ** 3 * n_it / (e0.elapsed_time(e1) / 1e3) / 1e12
        return {'tflops_sustained': round(sustained, 1), 'how': 'torch.matmul fp32 8192^3, allow_tf32=True, back to back for 2 s'}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


# ------------------------------------------------------------------------------------------------ CPU arm
def _oracle_trainer(cfg, arch_key):
    """CPU port of the reference path (oracle/): the same iteration on the host cores."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ref_step
    import torch_oracle as TO
    from architectures import network_architectures
    import mask_gen
    import inspect
    ctor = network_architectures.seg.get(cfg['arch'])
    kw = dict(pretrained=False) if 'pretrained' in inspect.signature(ctor).parameters else {}
    net = ctor(cfg['classes'], **kw)                                                        # shapes/keys only
    final = [k for k in net.state_dict() if ('layer5' in k or 'classifier.classifier.6' in k or 'final_clf' in k)
             and k.endswith('weight')]
    sd = TO.synth_state_dict(net.state_dict(), seed=0, logit_gain=40.0, final_keys=final)
    tr = ref_step.OracleMeanTeacher({'v3plus': 'deeplab3plus', 'v2': 'deeplab2', 'denseunet': 'denseunet'}[arch_key], sd, cfg['lr'])
    mg = mask_gen.BoxMaskGenerator(prop_range=0.5, n_boxes=1, random_aspect_ratio=True, prop_by_area=True,
                                   within_bounds=True, invert=True)
    return tr, mg


def cpu_step_time(args, batch, steps, warmup):
    from cutmix_semisup_seg_b200 import synthetic
    cfg = CFG[args.arch]
    # oneDNN convolutions stop scaling (and thrash) beyond a few dozen threads; state the count actually used
    torch.set_num_threads(max(1, min(32, os.cpu_count() or 1)))
    tr, mg = _oracle_trainer(cfg, args.arch)
    sup = synthetic.make_sup_batch(batch, cfg['h'], cfg['w'], cfg['classes'], 100)
    uns = synthetic.make_unsup_batch(batch, cfg['h'], cfg['w'], 200, mg, compact_masks=False)
    times = []
    for i in range(warmup + steps):
        t = time.time()
        tr.step(sup[0], sup[1], uns)
        times.append(time.time() - t)
    return times[warmup:], cfg


def cpu_baseline(args, sample_batch=1):
    """Reported baseline (not the target): one iteration of the oracle port on `sample_batch` images."""
    try:
        times, cfg = cpu_step_time(args, sample_batch, 1, 0)
    except Exception as e:  # keep the GPU line even if the host leg fails
        return {'value': None, 'unit': 'images/s', 'cores': torch.get_num_threads(), 'kind': 'port', 'sample': 'failed: %r' % (e,)}
    return {'value': round(sample_batch / times[0], 4), 'unit': 'images/s', 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': '1 iteration (4 fwd + 2 bwd + Adam + EMA) of the torch-CPU port of the reference path, batch {} at {}x{}'
                      .format(sample_batch, cfg['h'], cfg['w'])}


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    world = int(os.environ.get('WORLD_SIZE', 1))
    batch = 2          # train-mode BatchNorm of the DLv3+ pooling branch needs > 1 sample
    warm = min(args.warmup, 1)
    t0 = time.time()
    probe, cfg = cpu_step_time(args, batch, 1, 0)
    budget = 170.0
    steps = max(1, min(args.steps, int(budget / max(probe[0], 1e-3)) - warm))
    times, cfg = cpu_step_time(args, batch, steps, warm)
    sec = float(np.mean(times))
    val = batch / sec
    res = {'impl': 'reference', 'metric': 'images/sec', 'value': round(val, 4), 'unit': 'images/s', 'n_gpus': args.gpus,
           'steps': steps, 'warmup': warm, 'ms_per_step': round(sec * 1e3, 1), 'higher_is_better': True, 'scaling': 'weak',
           'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
           'config': {'workload': cfg['workload'], 'note': 'reference path (torch CPU port, oracle/) on host cores; each step is a '
                      'bounded sample of the workload: batch {} instead of {}'.format(batch, cfg['batch'])},
           'cpu_baseline': {'value': round(val, 4), 'unit': 'images/s', 'cores': torch.get_num_threads(), 'kind': 'port',
                            'sample': '{} iterations, batch {} at {}x{}'.format(steps, batch, cfg['h'], cfg['w'])},
           'e2e': {'value': round(val, 4), 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(res))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--arch', default='v3plus', choices=['v3plus', 'v2', 'denseunet'])
    ap.add_argument('--batch', type=int, default=0)
    ap.add_argument('--loss', default='cutmix', choices=['cutmix', 'ict', 'aug', 'vat'],
                    help='unsupervised branch: CutMix (the headline workload), ICT (train_seg_semisup_ict.py), augmentation '
                         'consistency (train_seg_semisup_aug_mt.py) or VAT (train_seg_semisup_vat_mt.py)')
    ap.add_argument('--u8-inputs', dest='u8_inputs', action='store_true',
                    help='also time the end-to-end loop with uint8 batches normalised on the device (extra key e2e_u8)')
    ap.add_argument('--eager', action='store_true', help='launch every kernel from Python instead of replaying CUDA graphs')
    ap.add_argument('--precision', default='tf32', choices=['tf32', '3xtf32'],
                    help='tensor-core mode of the timed iteration: single-pass TF32 (PyTorch/cuDNN default conv precision) or 3xTF32 '
                         '(the mode the tight parity tests run in); the other mode is timed as well on a few steps (precision_modes)')
    ap.add_argument('--no-second-precision', dest='no_second_precision', action='store_true',
                    help='skip the short run in the other precision mode')
    ap.add_argument('--no-tf32-peak', dest='measure_tf32_peak', action='store_false',
                    help='skip the 2 s cuBLAS tf32 matmul that places the tf32 library rate next to the roofline')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_b200(args)


if __name__ == '__main__':
    main()

"""-m gpu parity at the REAL geometry of the headline configuration (BASELINE.json configs[2]: DeepLab v3+ ResNet-101,
512 x 512, 19 classes, 16 images per GPU; and configs[1]: DeepLab v2, 321 x 321, N = 10).

  * the tensor-core convolution kernels at the exact shapes of the DeepLab v3+ head at 512 x 512 (ASPP 2048 -> 256 on a 64 x 64 map,
    dilation 12 / 24 / 36 -- every off-centre tap contributes --, the decoder's 304 -> 256 3x3 at 128 x 128, the HBM-bound
    256 -> 1024 1x1 with residual + ReLU over 32 images), single-pass TF32 and 3xTF32, against float64 F.conv2d;
  * the whole DeepLab v3+ (forward, every parameter gradient, BatchNorm running statistics) at 256 x 256 and 512 x 512 against the
    float64 oracle;
  * three CutMix mean-teacher iterations of DeepLab v3+ (frozen backbone BN, train-mode head BN, injected dropout masks, Adam,
    EMA) against the oracle's CPU iterations;
  * ONE full-size iteration (cfg2: N = 10 at 321 x 321; cfg3: N = 16 at 512 x 512) against the golden values produced by the
    UNMODIFIED reference modules (oracle/gen_golden_fullsize.py -> tests/golden/fullsize_*.npz), in 3xTF32 (the parity mode)
    and in single-pass TF32 (the throughput mode bench.py times), each with its tolerance written next to the assertion.

Float64 references of the large single-layer / whole-network cases are evaluated by torch ON THE GPU (ATen float64 kernels, no
TF32): the CPU would need minutes per case.  The iteration-level oracles run on the CPU like everywhere else in tests/."""
import json
import os
import sys
import warnings
from collections import OrderedDict

import numpy as np
import pytest
import torch
import torch.nn.functional as F

HERE = os.path.dirname(__file__)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'oracle'))
sys.path.insert(0, HERE)
import torch_oracle as TO  # noqa: E402
import ref_step  # noqa: E402
import mask_gen  # noqa: E402
import optim_weight_ema  # noqa: E402
import fullsize_recipe as R  # noqa: E402
from architectures import network_architectures as na  # noqa: E402

pytestmark = pytest.mark.gpu
dev = torch.device('cuda:0')


def _log(line):
    """Measured deviations are appended to $B200SEG_PARITY_LOG (a file under gpurun_out/) when set, so that DESIGN.md can quote
    them; the assertions below are what binds."""
    print(line)
    path = os.environ.get('B200SEG_PARITY_LOG')
    if path:
        with open(path, 'a') as f:
            f.write(line + '\n')


def relerr(got, ref):
    got = got.detach().double(); ref = ref.detach().double().to(got.device)
    return (got - ref).abs().max().item() / (ref.abs().max().item() + 1e-30)


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


@pytest.fixture(autouse=True)
def _exact_torch_reference():
    """torch's own GPU kernels serve as the float64 / float32 checker here: no TF32 in them."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


# ------------------------------------------------------------------------------------------ kernels at cfg3 geometry
CFG3_CONVS = [
    # N, H, W, Cin, Cout, k, dil          (DeepLab v3+ head at 512 x 512: SURVEY.md appendix A)
    (2, 64, 64, 2048, 256, 3, 12), (2, 64, 64, 2048, 256, 3, 24), (2, 64, 64, 2048, 256, 3, 36),
    (2, 64, 64, 2048, 256, 1, 1),              # ASPP 1x1 branch
    (2, 64, 64, 1280, 256, 1, 1),              # ASPP projection of the 5-branch concatenation
    (2, 128, 128, 304, 256, 3, 1),             # decoder: concat(48 + 256) -> 256
    (2, 128, 128, 256, 256, 3, 1),             # decoder second 3x3
    (3, 64, 64, 512, 512, 3, 4),               # layer4 3x3 d4
    (5, 64, 64, 256, 256, 3, 2),               # layer3 3x3 d2 (odd image count: phantom M tile of the CTA pair)
]


@pytest.mark.parametrize('case', CFG3_CONVS, ids=lambda c: 'x'.join(map(str, c)))
@pytest.mark.parametrize('n_split', [1, 3])
def test_conv_kernels_at_cfg3_geometry(case, n_split):
    """fprop / dgrad / wgrad through the engine's call path at the head's real shapes vs float64 F.conv2d.
    Tolerances relative to the output range: 2e-3 single-pass TF32 (10-bit mantissa products, fp32 accumulate);
    3xTF32: 5e-5 x max(1, K/4096) -- tcgen05 accumulates in fp32 with truncation, so the error grows with the reduction length
    K = Cin * k * k (18432 for the ASPP branches) -- and 1e-4 x max(1, pixels/8192) for wgrad's pixel reduction."""
    from cutmix_semisup_seg_b200.kernels import ActKernels
    from cutmix_semisup_seg_b200.acts import Act
    N, H, W, Cin, Cout, k, dil = case
    g = torch.Generator().manual_seed(sum(case))
    K = ActKernels(n_split=n_split)
    pad = dil * (k // 2)
    red = Cin * k * k
    tol = 2e-3 if n_split == 1 else 5e-5 * max(1.0, red / 4096.0)
    tol_w = 2e-3 if n_split == 1 else 1e-4 * max(1.0, N * H * W / 8192.0) ** 0.5
    x = torch.randn((N, Cin, H, W), generator=g)
    w = torch.randn((Cout, Cin, k, k), generator=g) / red ** 0.5
    dy = torch.randn((N, Cout, H, W), generator=g)
    xd = x.to(dev).double().requires_grad_(True)
    wd = w.to(dev).double().requires_grad_(True)
    y = F.conv2d(xd, wd, padding=pad, dilation=dil)
    y.backward(dy.to(dev).double())
    xa = Act(nhwc(x).to(dev), N, H, W, Cin)
    wk = w.permute(0, 2, 3, 1).contiguous().to(dev)
    out = Act.alloc(N, H, W, Cout, dev)
    K.conv_fwd(xa, wk, Cout, k, k, Cin, Cin, 1, pad, dil, out)
    e_f = relerr(out.to_nchw(), y)
    ga = Act(nhwc(dy).to(dev), N, H, W, Cout)
    dx = Act.alloc(N, H, W, Cin, dev)
    wt, ldb = K.transpose_w(wk, Cout, k * k, Cin)
    K.conv_dgrad(ga, wt, Cin, k, k, Cout, ldb, 1, pad, dil, dx)
    e_d = relerr(dx.to_nchw(), xd.grad)
    dw = torch.zeros(Cout, k * k, Cin, device=dev)
    K.conv_wgrad(ga, xa, dw, Cout, k, k, Cin, 1, pad, dil)
    e_w = relerr(dw.view(Cout, k, k, Cin), wd.grad.permute(0, 2, 3, 1))
    _log('conv cfg3 %s n_split=%d: fprop %.2e dgrad %.2e (tol %.1e) wgrad %.2e (tol %.1e)' % ('x'.join(map(str, case)), n_split,
                                                                                            e_f, e_d, tol, e_w, tol_w))
    assert e_f < tol and e_d < tol
    assert e_w < tol_w


@pytest.mark.parametrize('n_split', [1, 3])
def test_hbm_bound_1x1_with_residual_over_32_images(n_split):
    """layer3 conv3 of the batched trunk: 256 -> 1024 1x1 + folded BN scale / shift + residual + ReLU over 32 x 64 x 64 pixels
    (131072 GEMM rows; the PF build of the CTA-pair kernel), and its dgrad with fused addend + ReLU gate."""
    from cutmix_semisup_seg_b200.kernels import ActKernels
    from cutmix_semisup_seg_b200.acts import Act
    N, H, W, Cin, Cout = 32, 64, 64, 256, 1024
    g = torch.Generator().manual_seed(7)
    K = ActKernels(n_split=n_split)
    tol = 2e-3 if n_split == 1 else 5e-5
    x = torch.randn((N, Cin, H, W), generator=g)
    w = torch.randn((Cout, Cin, 1, 1), generator=g) / Cin ** 0.5
    scale = torch.rand((Cout,), generator=g) + 0.5
    shift = torch.randn((Cout,), generator=g)
    res = torch.randn((N, Cout, H, W), generator=g)
    xd, wd = x.to(dev).double(), w.to(dev).double()
    ref = torch.relu(F.conv2d(xd, wd) * scale.to(dev).double().view(1, -1, 1, 1) + shift.to(dev).double().view(1, -1, 1, 1)
                     + res.to(dev).double())
    xa = Act(nhwc(x).to(dev), N, H, W, Cin)
    wk = w.permute(0, 2, 3, 1).contiguous().to(dev)
    out = Act.alloc(N, H, W, Cout, dev)
    K.conv_fwd(xa, wk, Cout, 1, 1, Cin, Cin, 1, 0, 1, out, scale=scale.to(dev), shift=shift.to(dev),
               addend=Act(nhwc(res).to(dev), N, H, W, Cout), relu=True)
    e_f = relerr(out.to_nchw(), ref)
    del ref
    # backward of the same layer: dx = (dgrad(g) + partial) * (y_prev > 0), 1024 -> 256
    gy = torch.randn((N, Cout, H, W), generator=g)
    partial = torch.randn((N, Cin, H, W), generator=g)
    yprev = torch.randn((N, Cin, H, W), generator=g)
    xg = torch.zeros((N, Cin, H, W), device=dev, dtype=torch.double, requires_grad=True)
    F.conv2d(xg, wd).backward(gy.to(dev).double())
    refdx = (xg.grad + partial.to(dev).double()) * (yprev.to(dev) > 0)
    wt, ldb = K.transpose_w(wk, Cout, 1, Cin)
    dx = Act.alloc(N, H, W, Cin, dev)
    K.conv_dgrad(Act(nhwc(gy).to(dev), N, H, W, Cout), wt, Cin, 1, 1, Cout, ldb, 1, 0, 1, dx,
                 addend=Act(nhwc(partial).to(dev), N, H, W, Cin), gate=Act(nhwc(yprev).to(dev), N, H, W, Cin))
    e_d = relerr(dx.to_nchw(), refdx)
    _log('1x1 256->1024 residual N=32 n_split=%d: fprop %.2e dgrad %.2e (tol %.1e)' % (n_split, e_f, e_d, tol))
    assert e_f < tol and e_d < tol


# ------------------------------------------------------------------------------------------ whole network, large maps
def _dl3_compare(n, h, w, classes, precision, seed=1):
    """Full DeepLab v3+ (frozen backbone BN, train-mode head BN, injected dropout) vs the float64 oracle evaluated on the GPU."""
    net = na.seg.get('resnet101_deeplabv3plus_imagenet')(classes, pretrained=False)
    torch.manual_seed(seed)
    sd = TO.synth_state_dict(net.state_dict(), seed=seed)
    x = torch.randn(n, 3, h, w)
    dm = (torch.rand(n, -(-h // 8), -(-w // 8), 256) > 0.5).float()
    sd64 = OrderedDict((k, (v.double() if v.dtype == torch.float32 else v).clone().to(dev)) for k, v in sd.items())
    for k, p in net.named_parameters():
        if p.requires_grad:
            sd64[k].requires_grad_(True)
    yo = TO.deeplab3plus_forward(sd64, x.to(dev).double(), backbone_bn_train=False, head_bn_train=True,
                                 dropout_masks=[dm.permute(0, 3, 1, 2).to(dev).double()])
    dy = torch.randn(yo.shape)
    yo.backward(dy.to(dev).double())
    net.load_state_dict(sd)
    net.to(dev).train()
    net.freeze_batchnorm()
    net.b2_precision = precision
    for m in net.modules():
        if type(m).__name__ == 'B2Dropout':
            m.inject([dm])
    y = net(x.to(dev))
    y.backward(dy.to(dev))
    lerr = relerr(y, yo)
    errs = {}
    for k, p in net.named_parameters():
        gref = sd64[k].grad
        if not p.requires_grad or gref is None:
            continue
        errs[k] = relerr(p.grad, gref)
    stat = max((v.double() - sd64[k].detach()).abs().max().item() for k, v in net.state_dict().items() if 'running' in k)
    return lerr, errs, stat


@pytest.mark.parametrize('shape', [(3, 256, 256), (2, 512, 512)], ids=lambda s: 'x'.join(map(str, s)))
def test_deeplab3plus_large_maps_3xtf32(shape):
    """32 x 32 (dilation 12 and 24 in bounds) and 64 x 64 feature maps (all ASPP taps in bounds: the benchmark's geometry).
    Logits within 5e-4 of the float64 oracle's range (as for the small-crop test in test_gpu_nets.py; measured 1.4e-4 / 1.7e-4);
    parameter gradients follow the sqrt(forward error) law of ReLU networks (DESIGN.md 'Precision'): median over the 341
    tensors within 1e-1 (measured 2.7e-2), the worst head tensor (ASPP / decoder, downstream of the 33 residual units) within
    1e-1 (measured 4.8e-2)."""
    n, h, w = shape
    lerr, errs, stat = _dl3_compare(n, h, w, 19, '3xtf32')
    allv = sorted(errs.values())
    head = [v for k, v in errs.items() if 'classifier.classifier' in k]
    aspp = [v for k, v in errs.items() if 'classifier.aspp' in k]
    _log('dl3+ %s 3xtf32: logits %.2e, grads median %.2e max %.2e, decoder max %.2e, aspp max %.2e, running stats %.2e' % (
        shape, lerr, allv[len(allv) // 2], allv[-1], max(head), max(aspp), stat))
    assert lerr < 5e-4
    assert allv[len(allv) // 2] < 1e-1
    assert max(head) < 1e-1 and max(aspp) < 1e-1
    assert stat < 1e-3


def test_deeplab3plus_512_tf32_throughput_mode():
    """The precision bench.py times (single-pass TF32 = cuDNN's default) through all 101 layers on random weights: logits
    within 6e-2 of the float64 oracle's range at the benchmark's geometry (measured 3.5e-2), the median parameter gradient
    within 0.6 of its range (measured 0.33: sqrt(forward error) law -- sanity bounds; what this mode does to the LOSSES of real
    iterations is measured against the reference in test_fullsize_iterations_match_reference_golden[tf32-*])."""
    lerr, errs, stat = _dl3_compare(2, 512, 512, 19, 'tf32')
    allv = sorted(errs.values())
    _log('dl3+ (2,512,512) tf32: logits %.2e, grads median %.2e max %.2e, running stats %.2e' % (lerr, allv[len(allv) // 2], allv[-1], stat))
    assert lerr < 6e-2
    assert allv[len(allv) // 2] < 6e-1


# ------------------------------------------------------------------------------------------ iterations
def _trainer(cfg, precision, batch_trunk=True, fused_opt=True):
    from cutmix_semisup_seg_b200 import step as step_mod
    student = na.seg.get(cfg['kind'])(cfg['classes'], pretrained=False)
    sd = TO.synth_state_dict(student.state_dict(), seed=cfg['seed'], logit_gain=cfg['gain'], final_keys=R.final_keys(student.state_dict(), cfg))
    student.load_state_dict(sd)
    teacher = na.seg.get(cfg['kind'])(cfg['classes'], pretrained=False)
    student.to(dev); teacher.to(dev)
    student.b2_precision = teacher.b2_precision = precision
    for p in teacher.parameters():
        p.requires_grad = False
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        optim = step_mod.make_optimizer(student, 'adam', cfg['lr'], fused_kernel=fused_opt)
    ema = optim_weight_ema.EMAWeightOptimizer(teacher, student, 0.99)
    student.train(); teacher.train(); student.freeze_batchnorm(); teacher.freeze_batchnorm()
    mg = mask_gen.BoxMaskGenerator(0.5, invert=True)
    tr = step_mod.MeanTeacherStep(student, teacher, optim, ema, mg, conf_thresh=cfg['conf_thresh'], batch_trunk=batch_trunk)
    return tr, student, teacher, mg, sd


def _inject(student, teacher, dm):
    if dm is None:
        return
    for net, keys in ((student, ('sup', 'stu')), (teacher, ('tea0', 'tea1'))):
        drops = [m for m in net.modules() if type(m).__name__ == 'B2Dropout']
        assert len(drops) == 1
        drops[0].inject([dm[k] for k in keys])


def test_deeplab3plus_cutmix_iterations_match_oracle():
    """Three CutMix iterations of DeepLab v3+ at 2 x 256 x 256 (32 x 32 head map) vs the oracle's CPU iterations with the same
    dropout keep-masks.  Iteration 0 is pure forward / backward parity: both losses within 1e-4 relative (the north-star bound;
    measured 3.6e-6 / 3.5e-5), confidence rate within 1e-4 absolute (ONE pixel crossing the threshold is 7.6e-6 here).  From
    iteration 1 on each implementation has taken its own Adam steps: Adam's update is sign-like for near-zero gradients, so
    rounding-level gradient differences become +-lr weight differences and the runs drift apart at a rate set by the
    optimiser, not by the kernels (measured: sup 5.4e-5 -> 2.3e-4, cons 1.3e-4 -> 3.9e-4 over iterations 1 -> 2 with 131072
    pixels; at the full size of 4.2 M pixels all three iterations stay within 1e-4, see the golden test below).  Bounds for
    iterations >= 1: 1e-3 / 2e-3 relative, 5e-4 absolute."""
    cfg = dict(R.CONFIGS['cfg3_small'], conf_thresh=0.8)
    tr, student, teacher, mg, sd = _trainer(cfg, '3xtf32')
    orc = ref_step.OracleMeanTeacher('deeplab3plus', sd, cfg['lr'], conf_thresh=cfg['conf_thresh'])
    n, h, w = cfg['n'], cfg['h'], cfg['w']
    for it in range(3):
        (sx, sy), uns = R.batches(cfg, mg, compact_masks=True, it=it)
        dm = R.dropout_masks(cfg, it)
        _inject(student, teacher, dm)
        out = tr.step((sx.to(dev), sy.to(dev)), [{k: v.to(dev) for k, v in uns.items()}])
        uns_o = dict(uns)
        uns_o['mask_params'] = torch.from_numpy(TO.box_masks(uns['mask_params'].numpy(), (h, w), invert=True))
        drop = {k: v.permute(0, 3, 1, 2) for k, v in dm.items()}
        s_ref, c_ref, r_ref = orc.step(sx, sy, uns_o, drop={k: [v] for k, v in drop.items()})
        got = [float(out['sup_loss']), float(out['cons_loss']), float(out['conf_rate'])]
        _log('dl3+ cutmix iteration %d (2x256x256, 3xtf32): sup %.7f vs %.7f (rel %.1e), cons %.6e vs %.6e (rel %.1e), conf %.6f vs %.6f' % (
            it, got[0], s_ref, abs(got[0] - s_ref) / abs(s_ref), got[1], c_ref, abs(got[1] - c_ref) / (abs(c_ref) + 1e-30), got[2], r_ref))
        assert got[0] == pytest.approx(s_ref, rel=1e-4 if it == 0 else 1e-3)
        assert got[1] == pytest.approx(c_ref, rel=1e-4 if it == 0 else 2e-3, abs=1e-9)
        assert got[2] == pytest.approx(r_ref, abs=1e-4 if it == 0 else 5e-4)
    for name, net, ref in (('teacher', teacher, orc.teacher), ('student', student, orc.student)):
        worst = 0.0
        for k, v in net.state_dict().items():
            if v.dtype == torch.float32:
                r = ref[k].detach()
                worst = max(worst, (v.cpu() - r).abs().max().item() / (r.abs().max().item() + 1e-12))
        _log('dl3+ cutmix iterations: %s state max rel diff after 3 steps %.2e' % (name, worst))
        # Adam normalises gradients: a weight whose tiny gradient changes sign moves by up to +-lr per step, relative to a
        # tensor range that can be as small as ~1e-2 (BatchNorm shifts): 1e-3 of the range per step
        assert worst < 3e-3, (name, worst)


FULLSIZE_TOL = {
    # (precision, config) -> [(sup_loss rel, cons_loss rel, conf_rate abs) for iteration 0, the same for iterations >= 1]
    # 3xTF32 is the parity mode.  cfg3 (the headline configuration, 4.2 M pixels per batch) meets the north-star bound of 1e-4 on
    # every quantity in all three iterations (measured on a B200: sup <= 4.6e-5, cons <= 5.9e-5, conf <= 1.4e-5).
    ('3xtf32', 'cfg3'): [(1e-4, 1e-4, 1e-4), (1e-4, 1e-4, 1e-4)],
    # the 2-image stand-in: iteration 0 as tight; later iterations drift with the optimiser (see the oracle test above;
    # measured sup 9.4e-5, cons 5.5e-4)
    ('3xtf32', 'cfg3_small'): [(1e-4, 1e-4, 1e-4), (3e-4, 1.5e-3, 2e-4)],
    # cfg2 (DeepLab v2, no dropout): in iteration 0 teacher == student and the two views differ by 10 % noise, so the
    # consistency loss is a small difference of nearly equal probabilities (1.3e-3 against 0.3 later): its relative error is
    # cancellation-amplified and bounded at 5e-4 there
    ('3xtf32', 'cfg2'): [(1e-4, 5e-4, 1e-4), (3e-4, 1e-3, 2e-4)],
    # single-pass TF32 (what bench.py times; PyTorch's default convolution precision): measured on a B200 sup <= 1.1e-3 /
    # cons <= 2.3e-3 / conf <= 8.6e-4 (cfg3), 3.1e-3 / 1.8e-3 / 1.9e-3 (cfg3_small), 9.8e-3 / 2.9e-2 / 6.9e-3 (cfg2, whose
    # K = 18432 classifier feeds the soft-max directly)
    ('tf32', 'cfg3'): [(5e-3, 1e-2, 5e-3), (5e-3, 1e-2, 5e-3)],
    ('tf32', 'cfg3_small'): [(1e-2, 1e-2, 1e-2), (1e-2, 1e-2, 1e-2)],
    ('tf32', 'cfg2'): [(3e-2, 1e-1, 2e-2), (3e-2, 1e-1, 2e-2)],
}


@pytest.mark.parametrize('name', ['cfg3_small', 'cfg2', 'cfg3'])
@pytest.mark.parametrize('precision', ['3xtf32', 'tf32'])
def test_fullsize_iterations_match_reference_golden(name, precision):
    """Consecutive iterations (Adam moments, EMA teacher, BatchNorm buffers propagate) at BASELINE's full sizes vs the golden values
    computed by the unmodified reference modules on the CPU (fp32).  Tolerances: FULLSIZE_TOL above -- 3xTF32 is the parity mode
    (north star: losses within 1e-4 relative), single-pass TF32 is what bench.py times."""
    gold = np.load(os.path.join(HERE, 'golden', 'fullsize_%s.npz' % name))
    cfg = R.CONFIGS[name]
    tr, student, teacher, mg, sd = _trainer(cfg, precision)
    for it in range(cfg['iters']):
        t_sup, t_cons, t_conf = FULLSIZE_TOL[(precision, name)][min(it, 1)]
        (sx, sy), uns = R.batches(cfg, mg, compact_masks=True, it=it)
        _inject(student, teacher, R.dropout_masks(cfg, it))
        out = tr.step((sx.to(dev), sy.to(dev)), [{k: v.to(dev) for k, v in uns.items()}])
        got = [float(out['sup_loss']), float(out['cons_loss']), float(out['conf_rate'])]
        ref = [float(gold['sup_loss'][it]), float(gold['cons_loss'][it]), float(gold['conf_rate'][it])]
        _log('fullsize %s %s iteration %d: sup %.7f vs %.7f (rel %.1e), cons %.6e vs %.6e (rel %.1e), conf %.6f vs %.6f (abs %.1e)' % (
            name, precision, it, got[0], ref[0], abs(got[0] - ref[0]) / abs(ref[0]), got[1], ref[1], abs(got[1] - ref[1]) / abs(ref[1]),
            got[2], ref[2], abs(got[2] - ref[2])))
        assert got[0] == pytest.approx(ref[0], rel=t_sup)
        assert got[1] == pytest.approx(ref[1], rel=t_cons)
        assert got[2] == pytest.approx(ref[2], abs=t_conf)
    last = R.final_keys(student.state_dict(), cfg)[0]
    s_last = student.state_dict()[last].detach().cpu().numpy()[:, :64]        # the fixture keeps the first 64 input channels
    t_last = teacher.state_dict()[last].detach().cpu().numpy()[:, :64]
    rng = np.abs(gold['student_last']).max()
    e_s = np.abs(s_last - gold['student_last']).max() / rng
    e_t = np.abs(t_last - gold['teacher_last']).max() / rng
    _log('fullsize %s %s: final-layer weights after %d steps: student %.1e teacher %.1e of range' % (name, precision, cfg['iters'], e_s, e_t))
    # Adam moves every weight by at most lr per step (the sign of a tiny gradient may differ between implementations); the EMA
    # teacher takes 1 % of that per step
    assert e_s < 2.5 * cfg['iters'] * cfg['lr'] / rng + 1e-6
    assert e_t < 0.1 * cfg['iters'] * cfg['lr'] / rng + 1e-6

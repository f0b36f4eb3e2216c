"""-m gpu parity tests of the individual CUDA kernels (called through the C ABI) against the oracle / torch CPU
formulas.  Bit-exact where the arithmetic is elementwise fp32 (EMA, mix, masks); tolerances are written next to
each floating-point comparison."""
import json
import hashlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

HERE = os.path.dirname(__file__)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'oracle'))
import torch_oracle as TO  # noqa: E402
import mask_gen  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def be():
    from cutmix_semisup_seg_b200 import ops
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    return ops.default_backend()


dev = torch.device('cuda:0')


def relerr(got, ref):
    got = got.detach().cpu().double(); ref = ref.detach().cpu().double()
    return (got - ref).abs().max().item() / (ref.abs().max().item() + 1e-30)


# ------------------------------------------------------------------------------------------ bit-exact ops
@pytest.mark.parametrize('n', [1, 4097, 1000003])
def test_ema_flat_bit_exact(be, n):
    rs = np.random.RandomState(n)
    t = rs.randn(n).astype(np.float32); s = rs.randn(n).astype(np.float32)
    td, sd = torch.from_numpy(t).to(dev), torch.from_numpy(s).to(dev)
    be.ema_step_flat(td, sd, 0.99)
    assert np.array_equal(td.cpu().numpy(), TO.ema_update(t, s, 0.99))


def test_ema_optimizer_multi_tensor_bit_exact_incl_bn_buffers():
    import optim_weight_ema
    torch.manual_seed(0)

    def mk():
        return torch.nn.Sequential(torch.nn.Conv2d(3, 37, 3), torch.nn.BatchNorm2d(37), torch.nn.Conv2d(37, 5, 1)).to(dev)
    tea, stu = mk(), mk()
    for p in tea.parameters():
        p.requires_grad = False
    opt = optim_weight_ema.EMAWeightOptimizer(tea, stu, 0.99)
    with torch.no_grad():
        for p in stu.parameters():
            p.add_(torch.randn_like(p))
        stu[1].running_mean.normal_(); stu[1].running_var.uniform_(0.5, 2)
        stu[1].num_batches_tracked += 3
    ref = {k: v.cpu().numpy().copy() for k, v in tea.state_dict().items()}
    src = {k: v.cpu().numpy() for k, v in stu.state_dict().items()}
    for _ in range(3):
        opt.step()
        for k in ref:
            if ref[k].dtype == np.float32:
                ref[k] = TO.ema_update(ref[k], src[k], 0.99)
    for k, v in tea.state_dict().items():
        assert np.array_equal(v.cpu().numpy(), ref[k]), k
    assert int(tea[1].num_batches_tracked) == 0


def test_mix_and_cut_bit_exact(be):
    rs = np.random.RandomState(1)
    for shape in [(2, 3, 37, 41), (3, 1, 16, 16), (1, 3, 512, 512)]:
        n, c, h, w = shape
        a = rs.randn(*shape).astype(np.float32); b = rs.randn(*shape).astype(np.float32)
        m = (rs.rand(n, 1, h, w) > 0.5).astype(np.float32)
        m[0, 0, 0, :3] = [0.25, 0.5, 0.75]          # fractional mask values (valid-mask borders)
        a[0, 0, 1, 1] = -0.0; b[0, 0, 1, 2] = np.inf
        out = be.mix(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev), torch.from_numpy(m).to(dev))
        with np.errstate(invalid='ignore'):
            ref = TO.mix(a, b, m)
        got = out.cpu().numpy()
        assert np.array_equal(np.isnan(got), np.isnan(ref))              # inf * 0 -> NaN in both
        ok = ~np.isnan(ref)
        assert np.array_equal(got.view(np.uint32)[ok], ref.view(np.uint32)[ok])   # bit-exact incl. -0.0
        cut = be.mix(torch.from_numpy(a).to(dev), None, torch.from_numpy(m).to(dev))
        assert np.array_equal(cut.cpu().numpy().view(np.uint32), (a * m).astype(np.float32).view(np.uint32))


def test_box_masks_match_reference_golden():
    gold = json.load(open(os.path.join(HERE, 'golden', 'masks.json')))
    for case in gold:
        kw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in case['kwargs'].items()}
        gen = mask_gen.BoxMaskGenerator(**kw)
        shape = tuple(case['shape'])
        boxes = gen.generate_boxes(case['n'], shape, rng=np.random.RandomState(case['seed']))
        m = gen.torch_masks_from_params(torch.from_numpy(boxes), shape, dev)
        assert m.shape == (case['n'], 1) + shape and m.dtype == torch.float32
        assert hashlib.sha256(m.cpu().numpy().tobytes()).hexdigest() == case['sha256']


# ------------------------------------------------------------------------------------------ fused losses
@pytest.mark.parametrize('shape', [(2, 5, 6, 6), (2, 19, 33, 47), (3, 21, 40, 40), (2, 40, 9, 9)])
@pytest.mark.parametrize('fn', ['var', 'logits_var', 'logits_smoothl1', 'bce', 'kld'])
@pytest.mark.parametrize('pp', [False, True])
def test_consistency_kernel_vs_oracle(be, shape, fn, pp):
    torch.manual_seed(0)
    N, C, H, W = shape
    l0 = torch.randn(N, C, H, W) * 4; l1 = torch.randn(N, C, H, W) * 4; ls = torch.randn(N, C, H, W) * 4
    m = (torch.rand(N, 1, H, W) > 0.5).float(); um = torch.rand(N, 1, H, W)
    lsr = ls.clone().requires_grad_(True)
    loss, conf = TO.consistency_loss(l0, l1, lsr, m, um, fn, 0.6, pp)
    loss.backward()
    out4, dls = be.consistency(l0.to(dev), l1.to(dev), ls.to(dev), m.to(dev), um.to(dev), fn, 0.6, pp, 1.0, 1.0)
    o = out4.cpu()
    assert float(o[0]) == pytest.approx(float(loss), rel=2e-6)            # fp32 loss, double accumulation
    assert float(o[1]) == pytest.approx(float(conf), abs=1e-7)            # identical confidence decisions
    gtol = 5e-3 if fn == 'bce' else 1e-5        # bce gradient has 1/(p+1e-6) terms: fp32 conditioning
    assert relerr(dls.cpu() * o[2], lsr.grad) < gtol


def test_consistency_cut_mode_ramp_and_weight(be):
    torch.manual_seed(1)
    N, C, H, W = 2, 19, 20, 24
    lt = torch.randn(N, C, H, W) * 4; ls = torch.randn(N, C, H, W) * 4; lm = torch.rand(N, 1, H, W)
    lsr = ls.clone().requires_grad_(True)
    loss, conf = TO.consistency_loss(lt, None, lsr, None, lm, 'var', 0.5, False, ramp_val=0.3, rampup=5)
    (loss * 2.5).backward()
    out4, dls = be.consistency(lt.to(dev), None, ls.to(dev), None, lm.to(dev), 'var', 0.5, False, 0.3, 2.5)
    o = out4.cpu()
    assert float(o[0]) == pytest.approx(float(loss), rel=2e-6)
    assert float(o[3]) == pytest.approx(float(loss) * 2.5, rel=2e-6)
    assert relerr(dls.cpu() * o[2], lsr.grad) < 1e-5
    # conf_thresh <= 0 disables thresholding
    loss2, _ = TO.consistency_loss(lt, None, ls, None, lm, 'var', 0.0, False)
    out4, _ = be.consistency(lt.to(dev), None, ls.to(dev), None, lm.to(dev), 'var', 0.0, False, 1.0, 1.0)
    assert float(out4[0]) == pytest.approx(float(loss2), rel=2e-6)


def test_loss_block_survey_known_answers(be):
    gold = json.load(open(os.path.join(HERE, 'golden', 'loss_block.json')))
    torch.manual_seed(0)
    N, C, H, W = 2, 5, 6, 6
    l0 = torch.randn(N, C, H, W) * 4; l1 = torch.randn(N, C, H, W) * 4; ls = torch.randn(N, C, H, W) * 4
    um0 = torch.ones(N, 1, H, W); um1 = torch.ones(N, 1, H, W); um0[:, :, 0] = 0; um1[:, :, :, 0] = 0.5
    boxes = mask_gen.BoxMaskGenerator(0.5, invert=True).generate_boxes(N, (H, W), rng=np.random.RandomState(0))
    m = mask_gen.BoxMaskGenerator(0.5, invert=True).torch_masks_from_params(torch.from_numpy(boxes), (H, W), dev)
    um = be.mix(um0.to(dev), um1.to(dev), m)
    for key, exp in gold['cases'].items():
        fn, pp = key.rsplit('_pp', 1)
        out4, dls = be.consistency(l0.to(dev), l1.to(dev), ls.to(dev), m, um, fn, 0.6, bool(int(pp)), 1.0, 1.0)
        o = out4.cpu()
        assert float(o[0]) == pytest.approx(exp['loss'], rel=1e-5)
        assert float((dls * o[2].to(dev)).abs().sum()) == pytest.approx(exp['grad_l1'], rel=1e-4)
        assert float(o[1]) == pytest.approx(exp['conf_rate'], abs=1e-7)
    torch.manual_seed(1)
    lg = torch.randn(2, 5, 6, 6) * 2
    y = torch.randint(0, 5, (2, 1, 6, 6)); y[:, :, 0] = 255
    out3, dlg = be.cross_entropy(lg.to(dev), y[:, 0].contiguous().to(dev))
    assert float(out3[0]) == pytest.approx(gold['ce']['loss'], rel=1e-6)
    assert float(out3[1]) == gold['ce']['n_valid']
    assert float((dlg * out3[2]).abs().sum()) == pytest.approx(gold['ce']['grad_l1'], rel=1e-5)


@pytest.mark.parametrize('shape', [(2, 19, 33, 47), (1, 21, 64, 64), (2, 2, 8, 8)])
def test_cross_entropy_kernel(be, shape):
    torch.manual_seed(2)
    N, C, H, W = shape
    lg = torch.randn(N, C, H, W) * 3
    y = torch.randint(0, C, (N, H, W)); y[:, :2] = 255
    lgr = lg.clone().requires_grad_(True)
    ce = F.cross_entropy(lgr, y, ignore_index=255); ce.backward()
    out3, dlg = be.cross_entropy(lg.to(dev), y.to(dev))
    assert float(out3[0]) == pytest.approx(float(ce), rel=2e-6)
    assert relerr(dlg.cpu() * out3[2].cpu(), lgr.grad) < 1e-5
    assert float(dlg[:, :, :2].abs().max()) == 0.0           # ignored pixels get exactly zero gradient


# ------------------------------------------------------------------------------------------ tensor-core convs
def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


CONV_CASES = [
    # N, H, W, Cin, Cout, k, stride, dil
    (2, 16, 16, 64, 64, 1, 1, 1), (2, 16, 16, 64, 128, 3, 1, 1), (1, 64, 64, 256, 256, 3, 1, 2),
    (2, 13, 11, 304, 256, 3, 1, 1), (2, 32, 32, 128, 19, 1, 1, 1), (1, 24, 24, 512, 256, 3, 1, 12),
    (2, 16, 16, 128, 128, 3, 2, 1), (2, 17, 15, 64, 512, 1, 2, 1), (3, 5, 5, 64, 256, 3, 1, 1),
    (16, 1, 1, 2048, 256, 1, 1, 1), (2, 41, 41, 256, 1024, 1, 1, 1), (1, 21, 23, 2048, 21, 3, 1, 6),
]


@pytest.mark.parametrize('case', CONV_CASES, ids=lambda c: 'x'.join(map(str, c)))
@pytest.mark.parametrize('n_split', [1, 3])
def test_conv_fprop_dgrad_wgrad(case, n_split):
    """fprop / dgrad / wgrad of one convolution through the ActKernels interface (the engine's call path) vs
    float64 torch.  Tolerances (relative to the output range): 2e-3 single-pass TF32, 5e-5 3xTF32."""
    from cutmix_semisup_seg_b200.kernels import ActKernels
    from cutmix_semisup_seg_b200.acts import Act
    N, H, W, Cin, Cout, k, stride, dil = case
    torch.manual_seed(sum(case))
    K = ActKernels(n_split=n_split)
    # tcgen05 accumulates the fp32 sum with truncation: the 3xTF32 error grows with the reduction length
    tol = 2e-3 if n_split == 1 else 5e-5 * max(1.0, Cin * k * k / 4096.0)
    pad = dil * (k // 2)
    x = torch.randn(N, Cin, H, W, dtype=torch.double, requires_grad=True)
    w = (torch.randn(Cout, Cin, k, k, dtype=torch.double) / (Cin * k * k) ** 0.5).requires_grad_(True)
    y = F.conv2d(x, w, stride=stride, padding=pad, dilation=dil)
    dy = torch.randn(y.shape, dtype=torch.double)
    y.backward(dy)
    OH, OW = y.shape[2], y.shape[3]
    xa = Act(nhwc(x.detach().float()).to(dev), N, H, W, Cin)
    wd = w.detach().float().permute(0, 2, 3, 1).contiguous().to(dev)
    ldo = (Cout + 3) // 4 * 4
    out = Act.alloc(N, OH, OW, Cout, dev, ld=ldo)
    K.conv_fwd(xa, wd, Cout, k, k, Cin, Cin, stride, pad, dil, out)
    assert relerr(out.to_nchw(), y) < tol
    g = Act.alloc(N, OH, OW, Cout, dev, ld=ldo)
    g.base.zero_(); g.view4().copy_(nhwc(dy.float()).to(dev))
    dx = Act.alloc(N, H, W, Cin, dev)
    wt, ldb = K.transpose_w(wd, Cout, k * k, Cin)
    K.conv_dgrad(g, wt, Cin, k, k, Cout, ldb, stride, pad, dil, dx)
    assert relerr(dx.to_nchw(), x.grad) < tol
    dw = torch.zeros(Cout, k * k, Cin, device=dev)
    K.conv_wgrad(g, xa, dw, Cout, k, k, Cin, stride, pad, dil)
    assert relerr(dw.view(Cout, k, k, Cin), w.grad.permute(0, 2, 3, 1)) < tol
    # accumulate + row scale
    rs = torch.rand(Cout, device=dev) + 0.5
    K.conv_wgrad(g, xa, dw, Cout, k, k, Cin, stride, pad, dil, row_scale=rs, accumulate=True)
    ref = w.grad.permute(0, 2, 3, 1) * (1 + rs.cpu().double().view(-1, 1, 1, 1))
    assert relerr(dw.view(Cout, k, k, Cin), ref) < tol


@pytest.mark.parametrize('case', [
    # N, H, W, Cout(gemm K), Cin(gemm N = stats channels), k, dil, addend, sub
    (2, 12, 10, 48, 96, 3, 2, True, True),        # single-CTA kernel, ragged M tiles
    (3, 17, 12, 64, 256, 1, 1, True, False),      # CTA-pair kernel, 5 M tiles (phantom sixth)
    (2, 16, 16, 128, 512, 3, 1, False, True),     # CTA-pair kernel, two N tiles
    (1, 33, 41, 64, 64, 1, 1, False, False),
], ids=lambda c: 'x'.join(map(str, c)))
def test_conv_dgrad_fused_bn_statistics(case):
    """The dgrad epilogue that finishes a frozen-BN layer's output gradient also writes the column sums its affine
    parameters need; b2_bn_eval_param_grad_from_stats must reproduce the separate reduction (fp64 reference)."""
    from cutmix_semisup_seg_b200.kernels import ActKernels
    from cutmix_semisup_seg_b200.acts import Act
    N, H, W, Cout, Cin, k, dil, use_add, use_sub = case
    torch.manual_seed(sum(map(int, case)))
    K = ActKernels(n_split=3)
    pad = dil * (k // 2)
    w = torch.randn(Cout, Cin, k, k) / (Cin * k * k) ** 0.5
    g = torch.randn(N, Cout, H, W); partial = torch.randn(N, Cin, H, W); yprev = torch.randn(N, Cin, H, W)
    sub = torch.randn(N, Cin, H, W)
    gamma = torch.rand(Cin) + 0.5; beta = torch.randn(Cin)
    xg = torch.zeros(N, Cin, H, W, dtype=torch.double, requires_grad=True)
    F.conv2d(xg, w.double(), padding=pad, dilation=dil).backward(g.double())
    gin = (xg.grad + (partial.double() if use_add else 0)) * (yprev > 0)
    yv = yprev.double() - (sub.double() if use_sub else 0)
    sg = gin.sum(dim=(0, 2, 3)); sgy = (gin * yv).sum(dim=(0, 2, 3))
    ref_dbeta = sg; ref_dgamma = (sgy - beta.double() * sg) / gamma.double()
    wt, ldb = K.transpose_w(w.permute(0, 2, 3, 1).contiguous().to(dev), Cout, k * k, Cin)
    dx = Act.alloc(N, H, W, Cin, dev)
    st = K.conv_dgrad(Act(nhwc(g).to(dev), N, H, W, Cout), wt, Cin, k, k, Cout, ldb, 1, pad, dil, dx,
                      addend=Act(nhwc(partial).to(dev), N, H, W, Cin) if use_add else None,
                      gate=Act(nhwc(yprev).to(dev), N, H, W, Cin), want_stats=True,
                      stats_sub=Act(nhwc(sub).to(dev), N, H, W, Cin) if use_sub else None)
    assert relerr(dx.to_nchw(), gin) < 5e-5
    dgam = torch.full((Cin,), 3.0, device=dev); dbet = torch.full((Cin,), -2.0, device=dev)
    K.bn_eval_param_grad_from_stats(st, gamma.to(dev), beta.to(dev), dgam, dbet, False)
    assert relerr(dbet, ref_dbeta) < 1e-4 and relerr(dgam, ref_dgamma) < 1e-4
    K.bn_eval_param_grad_from_stats(st, gamma.to(dev), beta.to(dev), dgam, dbet, True)       # accumulate
    assert relerr(dbet, 2 * ref_dbeta) < 1e-4 and relerr(dgam, 2 * ref_dgamma) < 1e-4
    # and it agrees with the stand-alone reduction kernel on the same stored gradient
    dgam2 = torch.empty(Cin, device=dev); dbet2 = torch.empty(Cin, device=dev)
    K.bn_eval_param_grad(dx, Act(nhwc(yprev).to(dev), N, H, W, Cin), gamma.to(dev), beta.to(dev),
                         Act(nhwc(sub).to(dev), N, H, W, Cin) if use_sub else None, dgam2, dbet2, False)
    assert relerr(dgam2, ref_dgamma) < 1e-4 and relerr(dbet2, ref_dbeta) < 1e-4


@pytest.mark.parametrize('case', [
    # N, H, W, Cout(gemm K), Cin(gemm N), k, dil
    (2, 12, 10, 48, 96, 3, 2),         # single-CTA kernel, ragged M tiles
    (3, 17, 12, 64, 256, 1, 1),        # CTA-pair kernel (TMA epilogue), odd number of M tiles (phantom tile)
    (2, 16, 16, 128, 512, 3, 1),       # CTA-pair kernel, two N tiles
    (1, 64, 64, 64, 256, 3, 12),       # padding-only taps skipped per tile
], ids=lambda c: 'x'.join(map(str, c)))
def test_alternating_tile_direction_is_bit_identical(case):
    """Debug knob 12 (B200SEG_ALT_DIR): consecutive fprop / dgrad launches walk their tiles in opposite directions (L2 reuse of
    the tensor the previous launch wrote last).  The tile order must not change a single bit of the output or of the fused column
    statistics (their rows are indexed by the M tile, not by the time a tile is processed)."""
    from cutmix_semisup_seg_b200 import lib as L
    from cutmix_semisup_seg_b200.kernels import ActKernels
    from cutmix_semisup_seg_b200.acts import Act
    N, H, W, Cout, Cin, k, dil = case
    torch.manual_seed(sum(map(int, case)))
    K = ActKernels(n_split=1)
    pad = dil * (k // 2)
    w = (torch.randn(Cout, k, k, Cin) / (Cin * k * k) ** 0.5).to(dev)
    g = Act(torch.randn(N, H, W, Cout, device=dev), N, H, W, Cout)
    partial = Act(torch.randn(N, H, W, Cin, device=dev), N, H, W, Cin)
    yprev = Act(torch.randn(N, H, W, Cin, device=dev), N, H, W, Cin)
    wt, ldb = K.transpose_w(w, Cout, k * k, Cin)
    wf = (torch.randn(Cin, k, k, Cout) / (Cout * k * k) ** 0.5).to(dev)
    scale = (torch.rand(Cin) + 0.5).to(dev); shift = torch.randn(Cin).to(dev)

    def run():
        dx = Act.alloc(N, H, W, Cin, dev)
        st = K.conv_dgrad(g, wt, Cin, k, k, Cout, ldb, 1, pad, dil, dx, addend=partial, gate=yprev, want_stats=True)
        y = Act.alloc(N, H, W, Cin, dev)
        K.conv_fwd(g, wf, Cin, k, k, Cout, Cout, 1, pad, dil, y, scale=scale, shift=shift, addend=partial, relu=True)
        return dx.base.clone(), st[0].clone(), y.base.clone()

    lib = L.load()
    ref = run()                                   # knob off: every launch ascending
    lib.b2_debug_set(12, 1)
    try:
        a = run()                                 # dgrad ascending / fprop descending (or the other way round) ...
        K.conv_fwd(g, wf, Cin, k, k, Cout, Cout, 1, pad, dil, Act.alloc(N, H, W, Cin, dev))   # ... one launch shifts the phase ...
        b = run()                                 # ... so the second round runs each kernel in the other direction
    finally:
        lib.b2_debug_set(12, 0)
    for r, x, y in zip(ref, a, b):
        assert torch.equal(r, x) and torch.equal(r, y)


@pytest.mark.parametrize('case', [(8, 64, 64, 512, 256, 36), (2, 64, 64, 2048, 256, 36), (3, 41, 41, 512, 256, 12)])
def test_balanced_weight_gradient_plan_equals_one_unit_per_tap(case):
    """Dilated layers: the load-balanced plan of the weight-gradient kernel (more pixel splits, taps rotated from split to split,
    debug knob 14) must give the gradient of the plain plan up to the order of the fp32 slab sums, and both must match float64."""
    from cutmix_semisup_seg_b200 import lib as L
    from cutmix_semisup_seg_b200.kernels import ActKernels
    from cutmix_semisup_seg_b200.acts import Act
    N, H, W, Cin, Cout, dil = case
    torch.manual_seed(sum(case))
    K = ActKernels(n_split=1)
    x = torch.randn(N, H, W, Cin, device=dev); g = torch.randn(N, H, W, Cout, device=dev)
    xa, ga = Act(x, N, H, W, Cin), Act(g, N, H, W, Cout)
    lib = L.load()
    out = []
    for knob in (0, 1):
        lib.b2_debug_set(14, knob)
        try:
            dw = torch.zeros(Cout, 9, Cin, device=dev)
            K.conv_wgrad(ga, xa, dw, Cout, 3, 3, Cin, 1, dil, dil)
            out.append(dw)
        finally:
            lib.b2_debug_set(14, 1)
    scale = float(out[0].abs().max())
    assert float((out[0] - out[1]).abs().max()) < 2e-5 * scale
    # float64 reference on a sub-sample of the output channels (the full tensor is 1 TFLOP on the host)
    sel = torch.arange(0, Cout, 37)
    xd = x.permute(0, 3, 1, 2).double().cpu().requires_grad_(False)
    w = torch.zeros(len(sel), Cin, 3, 3, dtype=torch.double, requires_grad=True)
    y = F.conv2d(xd, w, padding=dil, dilation=dil)
    y.backward(g.permute(0, 3, 1, 2)[:, sel].double().cpu())
    ref = w.grad.permute(0, 2, 3, 1).reshape(len(sel), 9, Cin)
    assert relerr(out[1][sel], ref) < 2e-3                      # single-pass TF32


def test_conv_fused_epilogue_and_concat_slice():
    """scale/shift + residual + ReLU epilogue writing into a channel slice of a wider buffer; dgrad with the
    fused addend + ReLU gate (the backward fusion the engine relies on)."""
    from cutmix_semisup_seg_b200.kernels import ActKernels
    from cutmix_semisup_seg_b200.acts import Act
    torch.manual_seed(5)
    K = ActKernels(n_split=3)
    N, H, W, Cin, Cout = 2, 12, 10, 96, 48
    x = torch.randn(N, Cin, H, W); w = torch.randn(Cout, Cin, 3, 3) / (Cin * 9) ** 0.5
    scale = torch.rand(Cout) + 0.5; shift = torch.randn(Cout); res = torch.randn(N, Cout, H, W)
    ref = torch.relu(F.conv2d(x.double(), w.double(), padding=2, dilation=2) * scale.double().view(1, -1, 1, 1)
                     + shift.double().view(1, -1, 1, 1) + res.double())
    cat = Act.alloc(N, H, W, 112, dev); cat.base.fill_(7.0)
    sl = cat.slice(32, Cout)
    resa = Act(nhwc(res).to(dev), N, H, W, Cout)
    K.conv_fwd(Act(nhwc(x).to(dev), N, H, W, Cin), w.permute(0, 2, 3, 1).contiguous().to(dev), Cout, 3, 3, Cin, Cin, 1, 2, 2, sl,
               scale=scale.to(dev), shift=shift.to(dev), addend=resa, relu=True)
    assert relerr(sl.to_nchw(), ref) < 5e-5
    assert float(cat.base[..., :32].min()) == 7.0 and float(cat.base[..., 80:].min()) == 7.0     # neighbours untouched
    # dgrad: dx = (dgrad(g) + partial) gated by (y_prev > 0)
    g = torch.randn(N, Cout, H, W); partial = torch.randn(N, Cin, H, W); yprev = torch.randn(N, Cin, H, W)
    xg = torch.zeros(N, Cin, H, W, dtype=torch.double, requires_grad=True)
    F.conv2d(xg, w.double(), padding=2, dilation=2).backward(g.double())
    refdx = (xg.grad + partial.double()) * (yprev > 0)
    wt, ldb = K.transpose_w(w.permute(0, 2, 3, 1).contiguous().to(dev), Cout, 9, Cin)
    dx = Act.alloc(N, H, W, Cin, dev)
    K.conv_dgrad(Act(nhwc(g).to(dev), N, H, W, Cout), wt, Cin, 3, 3, Cout, ldb, 1, 2, 2, dx,
                 addend=Act(nhwc(partial).to(dev), N, H, W, Cin), gate=Act(nhwc(yprev).to(dev), N, H, W, Cin))
    assert relerr(dx.to_nchw(), refdx) < 5e-5


@pytest.mark.parametrize('shape', [(3, 19, 37, 53), (2, 21, 64, 64), (1, 2, 9, 1000)])
def test_argmax_confusion_matches_reference_evaluator(be, shape):
    """Fused argmax + confusion matrix (b2_argmax_confusion) vs the reference's numpy evaluator fed with torch.argmax:
    integer counts, so intersection / union / cm / mIoU must be identical (incl. ties, ignore label, NaN-free input)."""
    import evaluation
    n, c, h, w = shape
    g = torch.Generator().manual_seed(c)
    logits = torch.randn(n, c, h, w, generator=g)
    logits[:, :, :2] = torch.round(logits[:, :, :2])            # rows with many exact ties
    truth = torch.randint(0, c, (n, 1, h, w), generator=g)
    truth[:, :, -3:] = 255
    ref = evaluation.EvaluatorIoU(c)
    pred = torch.argmax(logits, dim=1).numpy()
    for i in range(n):
        ref.sample(truth[i, 0].numpy(), pred[i], ignore_value=255)
    got = evaluation.EvaluatorIoU(c)
    got.sample_logits(logits.to(dev), truth.to(dev), ignore_value=255)
    got.sample_logits(logits.to(dev), truth.to(dev), ignore_value=255)         # accumulates
    score = got.score()
    assert np.array_equal(got.cm, 2 * ref.cm)
    assert np.array_equal(got.intersection, 2 * ref.intersection) and np.array_equal(got.union, 2 * ref.union)
    assert np.array_equal(score, ref.score())
    cm = torch.zeros(c * c, dtype=torch.int64, device=dev)
    p = be.argmax_confusion(logits.to(dev), truth.to(dev), cm, ignore_value=255, want_pred=True)
    assert np.array_equal(p.cpu().numpy(), pred)

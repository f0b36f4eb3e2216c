"""-m gpu parity tests of the HBM-bound network operators (csrc/netops.cu, elementwise.cu) called through the engine's
kernel interface, against the torch-CPU float64 test double (tests/_emu_kernels.py = torch.nn.functional formulas).
Shapes cover the vectorised (C % 4 == 0, 16 B aligned) and the scalar fallbacks, channel slices (ld > C) and ragged
spatial sizes.  Tolerance 1e-5 of the output range (fp32 elementwise arithmetic), exact for max-pool and copies."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(__file__)
sys.path.insert(0, HERE)
from _emu_kernels import EmuKernels  # noqa: E402
from cutmix_semisup_seg_b200.acts import Act  # noqa: E402

pytestmark = pytest.mark.gpu
dev = torch.device('cuda:0')


@pytest.fixture(scope='module')
def K():
    from cutmix_semisup_seg_b200.kernels import ActKernels
    assert torch.cuda.is_available()
    return ActKernels(n_split=1)


E = EmuKernels()


def pair(n, h, w, c, ld=None, off=0, fill=None, seed=0):
    """(gpu Act, cpu Act) with identical contents; optional slice of a wider buffer."""
    g = torch.Generator().manual_seed(seed)
    ldv = (c + 3) // 4 * 4 if ld is None else ld
    base = torch.randn((n, h, w, ldv), generator=g) if fill is None else torch.full((n, h, w, ldv), float(fill))
    cpu = Act(base.clone(), n, h, w, c, ldv, off)
    gpu = Act(base.to(dev), n, h, w, c, ldv, off)
    return gpu, cpu


def rel(gpu_act, cpu_act):
    a = gpu_act.view4().cpu().double(); b = cpu_act.view4().double()
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-30)


def same_outside(gpu_act, cpu_act):
    """Bytes of the wider buffer outside the slice must be untouched (both started from the same contents)."""
    a = gpu_act.base.cpu().clone(); b = cpu_act.base.clone()
    a[..., gpu_act.off:gpu_act.off + gpu_act.c] = 0; b[..., cpu_act.off:cpu_act.off + cpu_act.c] = 0
    return torch.equal(a, b)


@pytest.mark.parametrize('shape', [(2, 3, 33, 41), (1, 3, 64, 64), (2, 4, 17, 9), (2, 3, 40, 300), (1, 3, 7, 129)])
def test_im2col_stem(K, shape):
    n, c, h, w = shape
    xg, xc = pair(n, h, w, c, ld=4, seed=1)
    oh, ow = (h + 6 - 7) // 2 + 1, (w + 6 - 7) // 2 + 1
    kpad = (49 * c + 31) // 32 * 32
    cg = K.im2col(xg, 7, 7, 2, 3, 1, oh, ow, kpad)
    cc = E.im2col(xc, 7, 7, 2, 3, 1, oh, ow, kpad)
    assert torch.equal(cg.base.cpu(), cc.base)


@pytest.mark.parametrize('shape', [(2, 33, 41, 64, False), (1, 32, 32, 64, False), (2, 17, 19, 6, False), (2, 33, 41, 64, True)])
def test_maxpool_fwd_bwd(K, shape):
    from cutmix_semisup_seg_b200 import engine
    n, h, w, c, ceil = shape
    xg, xc = pair(n, h, w, c, ld=c, seed=2)
    if ceil:
        oh, ow = -(-(h + 2 - 3) // 2) + 1, -(-(w + 2 - 3) // 2) + 1
        oh -= (oh - 1) * 2 >= h + 1; ow -= (ow - 1) * 2 >= w + 1
    else:
        oh, ow = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
    yg = Act.alloc(n, oh, ow, c, dev, ld=c); yc = Act.alloc(n, oh, ow, c, 'cpu', ld=c)
    ig = torch.empty((n, oh, ow, c), dtype=torch.uint8, device=dev); ic = torch.empty((n, oh, ow, c), dtype=torch.uint8)
    K.maxpool_fwd(xg, yg, ig); E.maxpool_fwd(xc, yc, ic)
    assert torch.equal(yg.base.cpu(), yc.base) and torch.equal(ig.cpu(), ic)
    dg, dc = pair(n, oh, ow, c, ld=c, seed=3)
    gx = Act.alloc(n, h, w, c, dev, ld=c); cx = Act.alloc(n, h, w, c, 'cpu', ld=c)
    K.maxpool_bwd(dg, ig, gx); E.maxpool_bwd(dc, ic, cx)
    assert rel(gx, cx) < 1e-6


@pytest.mark.parametrize('case', [
    # n, ih, iw, c, ld_in, oh, ow, align, out ld, out off
    (2, 16, 16, 256, 256, 32, 32, False, 304, 48),      # decoder: x2 into a concat slice
    (2, 9, 11, 21, 24, 65, 81, True, 24, 0),            # DeepLab v2: x8 align_corners=True, ragged channels
    (1, 7, 5, 6, 8, 13, 17, False, 8, 0),               # scalar path
    (2, 13, 17, 8, 8, 7, 5, False, 8, 0),               # down-sampling: some inputs receive no gradient
])
def test_bilinear_nhwc_fwd_bwd(K, case):
    n, ih, iw, c, ldi, oh, ow, align, ldo, off = case
    xg, xc = pair(n, ih, iw, c, ld=ldi, seed=4)
    yg, yc = pair(n, oh, ow, c, ld=ldo, off=off, fill=7.0)
    K.bilinear_fwd(xg, yg, align); E.bilinear_fwd(xc, yc, align)
    assert rel(yg, yc) < 1e-5 and same_outside(yg, yc)
    dg, dc = pair(n, oh, ow, c, ld=ldo, off=off, seed=5)
    gx, cx = pair(n, ih, iw, c, ld=ldi, seed=6)
    K.bilinear_bwd(dg, gx, align, accumulate=False); E.bilinear_bwd(dc, cx, align, accumulate=False)
    assert rel(gx, cx) < 1e-5
    K.bilinear_bwd(dg, gx, align, accumulate=True); E.bilinear_bwd(dc, cx, align, accumulate=True)
    assert rel(gx, cx) < 1e-5


@pytest.mark.parametrize('case', [(3, 64, 64, 256, 128, 128, False), (2, 9, 11, 64, 65, 81, True), (2, 33, 17, 48, 20, 9, False), (1, 5, 300, 8, 9, 517, True)])
def test_row_based_resize_and_im2col_equal_the_flat_kernels(K, case):
    """Debug knob 20: the row-based NHWC bilinear forward kernel / shared-memory stem im2col against their flat-index forms -- the
    same arithmetic, so every bit must agree."""
    from cutmix_semisup_seg_b200 import lib as L
    n, ih, iw, c, oh, ow, align = case
    torch.manual_seed(ih * iw + c)
    x = Act(torch.randn(n, ih, iw, c, device=dev), n, ih, iw, c)
    img = Act(torch.randn(n, 2 * oh + 5, 2 * ow + 3, 4, device=dev), n, 2 * oh + 5, 2 * ow + 3, 3, ld=4)
    soh, sow = (img.h + 6 - 7) // 2 + 1, (img.w + 6 - 7) // 2 + 1
    lib = L.load()
    outs = []
    for knob in (1, 0):
        lib.b2_debug_set(20, knob)
        try:
            y = Act.alloc(n, oh, ow, c, dev)
            K.bilinear_fwd(x, y, align)
            col = K.im2col(img, 7, 7, 2, 3, 1, soh, sow, 160)
            outs.append((y.base.clone(), col.base.clone()))
        finally:
            lib.b2_debug_set(20, 0)
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize('case', [(2, 16, 16, 19, 20, 64, 64, False), (2, 9, 11, 21, 24, 65, 81, True), (1, 6, 7, 5, 5, 11, 9, False),
                                  (1, 4, 4, 3, 4, 40, 44, True)])
def test_bilinear_nchw_fwd_bwd(K, case):
    """Final resize to NCHW logits and its backward from NCHW dlogits with the device-side gradient scale."""
    n, ih, iw, c, ldi, oh, ow, align = case
    xg, xc = pair(n, ih, iw, c, ld=ldi, seed=7)
    lg = torch.empty((n, c, oh, ow), device=dev); lc = torch.empty((n, c, oh, ow))
    K.bilinear_fwd_nchw(xg, lg, align); E.bilinear_fwd_nchw(xc, lc, align)
    assert (lg.cpu().double() - lc.double()).abs().max().item() <= 1e-5 * lc.abs().max().item()
    g = torch.randn((n, c, oh, ow), generator=torch.Generator().manual_seed(8))
    sc = torch.tensor([0.37])
    gx, cx = pair(n, ih, iw, c, ld=ldi, seed=9)
    K.bilinear_bwd_nchw(g.to(dev), gx, align, scale_dev=sc.to(dev), scale_host=2.0)
    E.bilinear_bwd_nchw(g, cx, align, scale_dev=sc, scale_host=2.0)
    assert rel(gx, cx) < 1e-5
    K.bilinear_bwd_nchw(g.to(dev), gx, align, accumulate=True); E.bilinear_bwd_nchw(g, cx, align, accumulate=True)
    assert rel(gx, cx) < 1e-5


@pytest.mark.parametrize('case', [(4, 9, 7, 64, 64, 0, True, True, True), (3, 5, 5, 48, 304, 0, True, False, False),
                                  (2, 6, 6, 10, 12, 0, False, True, False), (16, 1, 1, 256, 256, 0, True, False, False)])
def test_bn_train_stats_apply_bwd(K, case):
    n, h, w, c, ldo, off, relu, use_res, use_drop = case
    xg, xc = pair(n, h, w, c, seed=10)
    gam = torch.rand(c) + 0.5; bet = torch.randn(c)
    rm = torch.randn(c); rv = torch.rand(c) + 0.5
    mg, rg, rmg, rvg = torch.empty(c, device=dev), torch.empty(c, device=dev), rm.to(dev), rv.to(dev)
    mc, rc, rmc, rvc = torch.empty(c), torch.empty(c), rm.clone(), rv.clone()
    K.bn_stats(xg, 1e-5, 0.1, mg, rg, rmg, rvg); E.bn_stats(xc, 1e-5, 0.1, mc, rc, rmc, rvc)
    for a, b in ((mg, mc), (rg, rc), (rmg, rmc), (rvg, rvc)):
        assert (a.cpu().double() - b.double()).abs().max().item() <= 2e-5 * b.abs().max().item() + 1e-7
    mg.copy_(mc); rg.copy_(rc)                       # identical statistics for the element-wise parts
    resg, resc = pair(n, h, w, c, seed=11) if use_res else (None, None)
    mask = (torch.rand((n, h, w, c), generator=torch.Generator().manual_seed(12)) > 0.5).float() if use_drop else None
    yg, yc = pair(n, h, w, c, ld=ldo, off=off, fill=3.0)
    K.bn_apply(xg, mg, rg, gam.to(dev), bet.to(dev), relu, None if mask is None else mask.to(dev), 2.0, yg, residual=resg)
    E.bn_apply(xc, mc, rc, gam, bet, relu, mask, 2.0, yc, residual=resc)
    assert rel(yg, yc) < 1e-5 and same_outside(yg, yc)
    dg, dc = pair(n, h, w, c, seed=13)
    dxg = Act.alloc(n, h, w, c, dev); dxc = Act.alloc(n, h, w, c, 'cpu')
    gog = Act.alloc(n, h, w, c, dev) if use_res else None; goc = Act.alloc(n, h, w, c, 'cpu') if use_res else None
    dgam_g, dbet_g = torch.full((c,), 0.5, device=dev), torch.full((c,), -0.5, device=dev)
    dgam_c, dbet_c = torch.full((c,), 0.5), torch.full((c,), -0.5)
    yg.view4().copy_(yc.view4())                     # identical gate tensor
    K.bn_bwd(dg, xg, yg, mg, rg, gam.to(dev), relu, None if mask is None else mask.to(dev), 2.0, dxg, dgam_g, dbet_g, True, g_out=gog)
    E.bn_bwd(dc, xc, yc, mc, rc, gam, relu, mask, 2.0, dxc, dgam_c, dbet_c, True, g_out=goc)
    assert rel(dxg, dxc) < 2e-5
    assert (dgam_g.cpu().double() - dgam_c.double()).abs().max().item() <= 2e-5 * dgam_c.abs().max().item()
    assert (dbet_g.cpu().double() - dbet_c.double()).abs().max().item() <= 2e-5 * dbet_c.abs().max().item()
    if use_res:
        assert rel(gog, goc) < 1e-6


@pytest.mark.parametrize('case', [(4, 9, 7, 64, True), (3, 5, 5, 48, False), (2, 6, 6, 10, True), (16, 33, 31, 256, False)])
def test_bn_bwd_gate_recomputed_from_x_is_bit_identical(K, case):
    """b2_bn_bwd(gate_beta=beta): layers without a residual recompute the ReLU gate from the raw input with the instruction
    sequence of b2_bn_apply instead of reading the stored output.  Every result must equal the y-reading path bit for bit,
    including elements whose output is exactly zero or tiny (the sign test must see the value bn_apply stored)."""
    n, h, w, c, use_drop = case
    torch.manual_seed(n * 1000 + c)
    x = Act(torch.randn(n, h, w, c, device=dev) * 3.0, n, h, w, c)
    gam = (torch.rand(c) + 0.5).to(dev); bet = (torch.randn(c) * 0.5).to(dev)
    bet[::3] = 0.0                                   # zero shift: many outputs within rounding of zero
    mean = torch.empty(c, device=dev); rstd = torch.empty(c, device=dev)
    K.bn_stats(x, 1e-5, 0.1, mean, rstd, torch.zeros(c, device=dev), torch.ones(c, device=dev))
    x.view4()[0, 0, 0, :] = mean                     # exact zeros before the shift
    mask = (torch.rand(n, h, w, c, device=dev) > 0.5).float() if use_drop else None
    y = Act.alloc(n, h, w, c, dev)
    K.bn_apply(x, mean, rstd, gam, bet, True, mask, 2.0, y)
    dy = Act(torch.randn(n, h, w, c, device=dev), n, h, w, c)
    outs = []
    for gate_beta in (None, bet):
        dx = Act.alloc(n, h, w, c, dev)
        dgam = torch.full((c,), 0.25, device=dev); dbet = torch.full((c,), -0.75, device=dev)
        K.bn_bwd(dy, x, y, mean, rstd, gam, True, mask, 2.0, dx, dgam, dbet, True, gate_beta=gate_beta)
        outs.append((dx.view4().clone(), dgam, dbet))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    assert float(outs[0][0].abs().max()) > 0


@pytest.mark.parametrize('case', [(3, 8, 8, 2048, 2048), (2, 5, 7, 10, 12)])
def test_gap_and_broadcast(K, case):
    n, h, w, c, ld = case
    xg, xc = pair(n, h, w, c, ld=ld, seed=14)
    vg = Act.alloc(n, 1, 1, c, dev, ld=c); vc = Act.alloc(n, 1, 1, c, 'cpu', ld=c)      # pooled vectors are dense (N, C)
    K.gap_fwd(xg, vg); E.gap_fwd(xc, vc)
    assert rel(vg, vc) < 1e-5
    dxg, dxc = pair(n, h, w, c, ld=ld, seed=15)
    K.gap_bwd(vg, dxg, accumulate=True); E.gap_bwd(vc, dxc, accumulate=True)
    assert rel(dxg, dxc) < 1e-5
    K.gap_bwd(vg, dxg, accumulate=False); E.gap_bwd(vc, dxc, accumulate=False)
    assert rel(dxg, dxc) < 1e-5
    og, oc = pair(n, h, w, c, ld=ld + 8, off=4, fill=1.0)
    K.bcast_fwd(vg, og); E.bcast_fwd(vc, oc)
    assert rel(og, oc) < 1e-6 and same_outside(og, oc)
    sg = Act.alloc(n, 1, 1, c, dev, ld=c); sc = Act.alloc(n, 1, 1, c, 'cpu', ld=c)
    K.bcast_bwd(og, sg); E.bcast_bwd(oc, sc)
    assert rel(sg, sc) < 1e-5


@pytest.mark.parametrize('case', [(2, 9, 9, 64, 64, 0), (2, 9, 9, 48, 304, 48), (1, 5, 5, 7, 8, 0)])
def test_relu_gate_copy_colsum(K, case):
    n, h, w, c, ld, off = case
    gg, gc = pair(n, h, w, c, ld=ld, off=off, seed=16)
    yg, yc = pair(n, h, w, c, seed=17)
    K.relu_gate(gg, yg); E.relu_gate(gc, yc)
    assert torch.equal(gg.base.cpu(), gc.base)
    dg, dc = pair(n, h, w, c, seed=18)
    K.copy_act(dg, gg); E.copy_act(dc, gc)
    assert torch.equal(dg.view4().cpu(), dc.view4())
    K.copy_act(gg, dg, accumulate=True); E.copy_act(gc, dc, accumulate=True)
    assert torch.equal(gg.base.cpu(), gc.base)
    og, oc = torch.full((c,), 2.0, device=dev), torch.full((c,), 2.0)
    K.colsum(gg, og, True); E.colsum(gc, oc, True)
    assert (og.cpu().double() - oc.double()).abs().max().item() <= 1e-5 * oc.abs().max().item()


class _BN(object):
    pass


@pytest.mark.parametrize('case', [(2, 16, 16, 64, 64 * 9), (1, 33, 41, 64, 147), (3, 9, 9, 256, 1024)])
def test_frozen_bn_param_grads_from_weight_gradient(K, case):
    """b2_bn_eval_param_grad_wdot: dbeta (+)= column sums of g, dgamma = <W, dW>/gamma - invstd*mean*dbeta; compared
    with the same formula in float64 (the algebra itself is checked against autograd in tests/test_engine_emu.py)."""
    n, h, w, c, row = case
    gg, gc = pair(n, h, w, c, seed=19)
    gen = torch.Generator().manual_seed(20)
    W = torch.randn((c, row), generator=gen); gW = torch.randn((c, row), generator=gen)
    bn = _BN(); bn.eps = 1e-5
    bn.weight = torch.rand(c, generator=gen) + 0.5; bn.running_mean = torch.randn(c, generator=gen)
    bn.running_var = torch.rand(c, generator=gen) + 0.5
    bng = _BN(); bng.eps = bn.eps
    bng.weight, bng.running_mean, bng.running_var = bn.weight.to(dev), bn.running_mean.to(dev), bn.running_var.to(dev)
    for acc in (False, True):
        dgam_g, dbet_g = torch.full((c,), 9.0, device=dev), torch.full((c,), 0.25, device=dev)
        dgam_c, dbet_c = torch.full((c,), 9.0), torch.full((c,), 0.25)
        K.bn_eval_param_grad_wdot(None, gg, W.to(dev), gW.to(dev), bng, dgam_g, dbet_g, acc)
        E.bn_eval_param_grad_wdot(None, gc, W, gW, bn, dgam_c, dbet_c, acc)
        assert (dbet_g.cpu().double() - dbet_c.double()).abs().max().item() <= 1e-5 * dbet_c.abs().max().item()
        assert (dgam_g.cpu().double() - dgam_c.double()).abs().max().item() <= 1e-5 * dgam_c.abs().max().item()

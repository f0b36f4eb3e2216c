"""Diagnostic sweep of the CUDA kernels against torch-CPU formulas (development tool; the graded
parity tests live in tests/).  Prints one line per case so a failing encoding is easy to spot.

usage: python tools/gpu_probe.py <group> [...]   groups: elem loss conv dgrad wgrad
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F

from cutmix_semisup_seg_b200 import ops as O
from cutmix_semisup_seg_b200 import lib as L

dev = torch.device('cuda:0')
be = O.CudaBackend()


def report(name, got, ref, tol):
    got = got.detach().cpu().double(); ref = ref.detach().cpu().double()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-30
    ok = err <= tol * scale
    print('{:<70s} maxerr {:.3e} rel {:.3e} {}'.format(name, err, err / scale, 'OK' if ok else 'FAIL'), flush=True)
    return ok


def g_elem():
    torch.manual_seed(0)
    t = torch.randn(1000003); s = torch.randn(1000003)
    ref = t.clone(); ref.mul_(0.99); ref.add_(s * (1.0 - 0.99))
    td = t.to(dev); sd = s.to(dev)
    be.ema_step_flat(td, sd, 0.99)
    print('ema flat bit-exact:', torch.equal(td.cpu(), ref))
    a = torch.randn(2, 3, 37, 41); b = torch.randn(2, 3, 37, 41); m = (torch.rand(2, 1, 37, 41) > 0.5).float()
    ref = a * (1 - m) + b * m
    out = be.mix(a.to(dev), b.to(dev), m.to(dev))
    print('mix bit-exact:', torch.equal(out.cpu(), ref))
    out = be.mix(a.to(dev), None, m.to(dev))
    print('cut bit-exact:', torch.equal(out.cpu(), a * m))
    boxes = torch.tensor([[[1, 6, 1, 7]], [[1, 6, 2, 8]]], dtype=torch.int32, device=dev)
    mk = be.box_mask_rasterize(boxes, 8, 8, 0.0)
    print('mask sums', mk.sum(dim=(1, 2, 3)).tolist())
    x = torch.randn(5000, device=dev)
    hi, lo = be.split_tf32(x)
    print('split exact:', torch.equal((hi + lo).cpu(), x.cpu()), 'hi low bits zero:',
          int((hi.view(torch.int32) & 0x1fff).abs().max().item()) == 0)
    w = torch.randn(70, 9, 50, device=dev)
    wt = be.transpose_w(w, 70, 9, 50)
    print('transpose_w:', torch.equal(wt.cpu(), w.permute(2, 1, 0).contiguous().cpu()))


def ref_consistency(l0, l1, ls, m, lmask, loss_fn, tau, per_pixel):
    ls = ls.clone().requires_grad_(True)
    lt = l0 * (1 - m) + l1 * m if l1 is not None else l0
    pt = F.softmax(lt, dim=1); ps = F.softmax(ls, dim=1)
    loss_mask = lmask
    conf_rate = torch.tensor(1.0)
    if tau > 0:
        conf = (pt.max(dim=1)[0] >= tau).float()[:, None]
        conf_rate = conf.mean()
        loss_mask = loss_mask * (conf if per_pixel else conf.mean())
    C = ls.shape[1]
    if loss_fn == 'var':
        q = ((ps - pt) ** 2).sum(dim=1, keepdim=True)
    elif loss_fn == 'logits_var':
        q = ((ls - lt) ** 2).sum(dim=1, keepdim=True) / (C ** 0.5)
    elif loss_fn == 'logits_smoothl1':
        q = F.smooth_l1_loss(ls, lt, reduction='none').sum(dim=1, keepdim=True) / (C ** 0.5)
    elif loss_fn == 'bce':
        eps = 1e-6
        q = (-(pt * torch.log(ps + eps) + (1 - pt) * torch.log(1 - ps + eps))).sum(dim=1, keepdim=True)
    else:
        q = F.kl_div(F.log_softmax(ls, dim=1), pt, reduction='none').sum(dim=1, keepdim=True)
    loss = (q * loss_mask).mean()
    loss.backward()
    return loss.detach(), conf_rate, ls.grad


def g_loss():
    torch.manual_seed(0)
    for (N, C, H, W) in [(2, 5, 6, 6), (2, 19, 33, 47), (3, 21, 40, 40), (2, 40, 9, 9)]:
        l0 = torch.randn(N, C, H, W) * 4; l1 = torch.randn(N, C, H, W) * 4; ls = torch.randn(N, C, H, W) * 4
        m = (torch.rand(N, 1, H, W) > 0.5).float()
        um = torch.rand(N, 1, H, W)
        for fn in ['var', 'logits_var', 'logits_smoothl1', 'bce', 'kld']:
            for pp in [False, True]:
                loss, cr, g = ref_consistency(l0, l1, ls, m, um, fn, 0.6, pp)
                out4, dls = be.consistency(l0.to(dev), l1.to(dev), ls.to(dev), m.to(dev), um.to(dev), fn, 0.6, pp, 1.0, 1.0)
                o = out4.cpu()
                report('cons {} {} pp={} loss'.format((N, C, H, W), fn, pp), o[0], loss, 2e-6)
                report('   conf_rate', o[1], cr, 1e-7)
                report('   grad', dls.cpu() * o[2], g, 1e-5)
        lg = torch.randn(N, C, H, W) * 2
        y = torch.randint(0, C, (N, H, W)); y[:, 0] = 255
        lgr = lg.clone().requires_grad_(True)
        ce = F.cross_entropy(lgr, y, ignore_index=255); ce.backward()
        out3, dlg = be.cross_entropy(lg.to(dev), y.to(dev))
        o = out3.cpu()
        report('ce {} loss'.format((N, C, H, W)), o[0], ce.detach(), 2e-6)
        report('   grad', dlg.cpu() * o[2], lgr.grad, 1e-5)


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def run_fprop(N, H, W, Cin, Cout, k, stride, dil, n_split=1, epi=False, flat=False):
    pad = dil * (k // 2)
    x = torch.randn(N, Cin, H, W); w = torch.randn(Cout, Cin, k, k) / (Cin * k * k) ** 0.5
    ref = F.conv2d(x.double(), w.double(), stride=stride, padding=pad, dilation=dil)
    OH, OW = ref.shape[2], ref.shape[3]
    xd = nhwc(x).to(dev); wd = w.permute(0, 2, 3, 1).contiguous().to(dev)   # KRSC
    ldd = ((Cout + 3) // 4) * 4
    out = torch.zeros(N, OH, OW, ldd, device=dev)
    kw = {}
    if epi:
        scale = torch.rand(Cout) + 0.5; shift = torch.randn(Cout)
        add = torch.randn(N, OH, OW, ldd)
        ref = torch.relu(ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1) + add[..., :Cout].permute(0, 3, 1, 2).double())
        kw = dict(scale=scale.to(dev), shift=shift.to(dev), relu=True)
        addd = add.to(dev)
        kw['addend'] = addd.data_ptr(); kw['ld_add'] = ldd
    a_lo = b_lo = None
    xa, wb = xd, wd
    if n_split > 1:
        xa, xlo = be.split_tf32(xd); wb, wlo = be.split_tf32(wd)
        a_lo, b_lo = xlo.data_ptr(), wlo.data_ptr()
    taps = O.conv_taps(k, k, dil, pad)
    if flat:
        assert k == 1 and stride == 1
        be.conv_gemm(xa.data_ptr(), 1, 1, N * H * W, Cin, Cin, wb.data_ptr(), Cout, 1, Cin, out.data_ptr(), 1, N * H * W, 1, N * H * W,
                     ldd, taps, a_lo_ptr=a_lo, b_lo_ptr=b_lo, n_split=n_split, **kw)
    else:
        be.conv_gemm(xa.data_ptr(), N, H, W, Cin, Cin, wb.data_ptr(), Cout, k * k, Cin, out.data_ptr(), OH, OW, OH, OW, ldd, taps,
                     istride=stride, a_lo_ptr=a_lo, b_lo_ptr=b_lo, n_split=n_split, **kw)
    torch.cuda.synchronize()
    got = out[..., :Cout].permute(0, 3, 1, 2)
    tol = 2e-3 if n_split == 1 else 5e-5
    return report('fprop N{} {}x{} {}->{} k{} s{} d{} split{} epi{} flat{}'.format(N, H, W, Cin, Cout, k, stride, dil, n_split, int(epi), int(flat)),
                  got, ref, tol)


def g_conv():
    torch.manual_seed(1)
    run_fprop(2, 16, 16, 64, 64, 1, 1, 1)
    run_fprop(2, 16, 16, 64, 64, 1, 1, 1, flat=True)
    run_fprop(2, 16, 16, 64, 64, 1, 1, 1, n_split=3)
    run_fprop(2, 16, 16, 64, 64, 1, 1, 1, n_split=4)
    run_fprop(2, 16, 16, 64, 128, 3, 1, 1)
    run_fprop(2, 16, 16, 64, 128, 3, 1, 1, n_split=3)
    run_fprop(1, 64, 64, 256, 256, 3, 1, 2)
    run_fprop(2, 13, 13, 96, 256, 3, 1, 2, epi=True)
    run_fprop(2, 13, 11, 304, 256, 3, 1, 1, n_split=3)
    run_fprop(2, 32, 32, 128, 19, 1, 1, 1, n_split=3)
    run_fprop(2, 32, 32, 128, 48, 1, 1, 1, epi=True)
    run_fprop(1, 24, 24, 512, 512, 3, 1, 12, n_split=3)
    run_fprop(2, 16, 16, 128, 128, 3, 2, 1, n_split=3)
    run_fprop(2, 17, 15, 64, 512, 1, 2, 1, n_split=3)
    run_fprop(3, 5, 5, 64, 256, 3, 1, 1, n_split=3)
    run_fprop(16, 1, 1, 2048, 256, 1, 1, 1, n_split=3)
    run_fprop(2, 41, 41, 256, 1024, 1, 1, 1, flat=True)
    run_fprop(2, 64, 64, 160, 64, 1, 1, 1, flat=True, n_split=3)


def run_dgrad(N, H, W, Cin, Cout, k, dil, n_split=3):
    pad = dil * (k // 2)
    x = torch.randn(N, Cin, H, W, dtype=torch.double, requires_grad=True)
    w = (torch.randn(Cout, Cin, k, k) / (Cin * k * k) ** 0.5)
    y = F.conv2d(x, w.double(), padding=pad, dilation=dil)
    dy = torch.randn(y.shape)
    y.backward(dy.double())
    ref = x.grad
    dyd = nhwc(dy).to(dev)
    wd = w.permute(0, 2, 3, 1).contiguous().to(dev)       # (Cout, T, Cin)
    wt = be.transpose_w(wd, Cout, k * k, Cin)              # (Cin, T, Cout)
    out = torch.zeros(N, H, W, Cin, device=dev)
    a_lo = b_lo = None
    ya, wb = dyd, wt
    if n_split > 1:
        ya, ylo = be.split_tf32(dyd); wb, wlo = be.split_tf32(wt)
        a_lo, b_lo = ylo.data_ptr(), wlo.data_ptr()
    be.conv_gemm(ya.data_ptr(), N, H, W, Cout, Cout, wb.data_ptr(), Cin, k * k, Cout, out.data_ptr(), H, W, H, W, Cin,
                 O.dgrad_taps(k, k, dil, pad), a_lo_ptr=a_lo, b_lo_ptr=b_lo, n_split=n_split)
    torch.cuda.synchronize()
    report('dgrad N{} {}x{} {}->{} k{} d{} split{}'.format(N, H, W, Cin, Cout, k, dil, n_split), out.permute(0, 3, 1, 2), ref,
           2e-3 if n_split == 1 else 5e-5)


def g_dgrad():
    torch.manual_seed(2)
    run_dgrad(2, 16, 16, 64, 128, 3, 1)
    run_dgrad(2, 13, 13, 96, 256, 3, 2)
    run_dgrad(2, 16, 16, 256, 64, 1, 1)
    run_dgrad(1, 24, 24, 128, 64, 3, 12)
    run_dgrad(2, 16, 16, 64, 128, 3, 1, n_split=1)


def run_wgrad(N, H, W, Cin, Cout, k, stride, dil, n_split=3, variant=0, max_ctas=0):
    pad = dil * (k // 2)
    x = torch.randn(N, Cin, H, W)
    w = torch.zeros(Cout, Cin, k, k, dtype=torch.double, requires_grad=True)
    y = F.conv2d(x.double(), w, stride=stride, padding=pad, dilation=dil)
    dy = torch.randn(y.shape)
    y.backward(dy.double())
    ref = w.grad.permute(0, 2, 3, 1).contiguous()           # (Cout, R, S, Cin)
    OH, OW = y.shape[2], y.shape[3]
    xd = nhwc(x).to(dev); dyd = nhwc(dy).to(dev)
    dw = torch.zeros(Cout, k * k, Cin, device=dev)
    y_lo = x_lo = None
    ya, xa = dyd, xd
    if n_split > 1:
        ya, ylo = be.split_tf32(dyd); xa, xlo = be.split_tf32(xd)
        y_lo, x_lo = ylo.data_ptr(), xlo.data_ptr()
    L.load().b2_debug_set(1, variant)
    be.conv_wgrad(ya.data_ptr(), N, OH, OW, Cout, Cout, xa.data_ptr(), H, W, Cin, Cin, dw.data_ptr(), O.conv_taps(k, k, dil, pad), k * k,
                  istride=stride, dy_lo_ptr=y_lo, x_lo_ptr=x_lo, n_split=n_split, max_ctas=max_ctas, device=dev)
    torch.cuda.synchronize()
    L.load().b2_debug_set(1, 0)
    report('wgrad N{} {}x{} {}->{} k{} s{} d{} split{} var{} ctas{}'.format(N, H, W, Cin, Cout, k, stride, dil, n_split, variant, max_ctas),
           dw.view(Cout, k, k, Cin), ref, 2e-3 if n_split == 1 else 5e-5)


def g_wgrad():
    torch.manual_seed(3)
    for variant in (0, 1):
        run_wgrad(2, 16, 16, 64, 64, 1, 1, 1, variant=variant)
        run_wgrad(2, 16, 16, 64, 128, 3, 1, 1, variant=variant, max_ctas=4)
    run_wgrad(2, 16, 16, 64, 128, 3, 1, 1)
    run_wgrad(2, 13, 13, 96, 256, 3, 1, 2)
    run_wgrad(2, 13, 11, 304, 256, 3, 1, 1)
    run_wgrad(2, 32, 32, 128, 20, 1, 1, 1)
    run_wgrad(2, 32, 32, 48, 48, 1, 1, 1)
    run_wgrad(1, 24, 24, 64, 64, 3, 1, 12)
    run_wgrad(2, 16, 16, 128, 128, 3, 2, 1)
    run_wgrad(2, 17, 15, 64, 512, 1, 2, 1)
    run_wgrad(16, 1, 1, 2048, 256, 1, 1, 1)
    run_wgrad(2, 16, 16, 64, 128, 3, 1, 1, n_split=1)




def g_netops():
    torch.manual_seed(4)
    N, C, H, W = 2, 64, 21, 17
    x = torch.randn(N, C, H, W); xh = nhwc(x).to(dev)
    # max pool floor / ceil
    for ceil in (False, True):
        xr = x.clone().requires_grad_(True)
        y = F.max_pool2d(xr, 3, 2, 1, ceil_mode=ceil)
        dy = torch.randn(y.shape); y.backward(dy)
        OH, OW = y.shape[2], y.shape[3]
        yd = torch.empty(N, OH, OW, C, device=dev); idx = torch.empty(N, OH, OW, C, device=dev, dtype=torch.uint8)
        be.maxpool_fwd(xh.data_ptr(), yd.data_ptr(), idx.data_ptr(), N, H, W, C, OH, OW)
        report('maxpool ceil={} fwd'.format(ceil), yd.permute(0, 3, 1, 2), y.detach(), 0)
        dx = torch.empty(N, H, W, C, device=dev)
        be.maxpool_bwd(nhwc(dy).to(dev).data_ptr(), idx.data_ptr(), dx.data_ptr(), N, H, W, C, OH, OW)
        report('maxpool ceil={} bwd'.format(ceil), dx.permute(0, 3, 1, 2), xr.grad, 1e-6)
    # bilinear
    for (ih, iw, oh, ow, ac, c) in [(6, 5, 41, 33, True, 21), (8, 8, 16, 16, False, 256), (16, 12, 64, 48, False, 19), (1, 1, 7, 9, False, 8), (9, 9, 9, 9, True, 5)]:
        xs = torch.randn(2, c, ih, iw, requires_grad=True)
        y = F.interpolate(xs, size=(oh, ow), mode='bilinear', align_corners=ac)
        dy = torch.randn(y.shape); y.backward(dy)
        xsd = nhwc(xs.detach()).to(dev)
        for to_nchw in (False, True):
            if to_nchw:
                yd = torch.empty(2, c, oh, ow, device=dev)
                be.bilinear_fwd(xsd.data_ptr(), yd.data_ptr(), 2, ih, iw, c, c, oh, ow, c, ac, True)
                got = yd
                dyd = dy.to(dev)
            else:
                yd = torch.empty(2, oh, ow, c, device=dev)
                be.bilinear_fwd(xsd.data_ptr(), yd.data_ptr(), 2, ih, iw, c, c, oh, ow, c, ac, False)
                got = yd.permute(0, 3, 1, 2)
                dyd = nhwc(dy).to(dev)
            report('bilinear {}x{}->{}x{} ac={} nchw={} fwd'.format(ih, iw, oh, ow, ac, to_nchw), got, y.detach(), 2e-6)
            dx = torch.empty(2, ih, iw, c, device=dev)
            be.bilinear_bwd(dyd.data_ptr(), dx.data_ptr(), 2, ih, iw, c, c, oh, ow, c, ac, to_nchw)
            report('   bwd', dx.permute(0, 3, 1, 2), xs.grad, 5e-6)
    # gap / bcast
    y = torch.empty(N, C, device=dev)
    be.gap_fwd(xh.data_ptr(), y.data_ptr(), N, H * W, C, C)
    report('gap fwd', y, x.mean(dim=(2, 3)), 1e-6)
    v = torch.randn(N, C, device=dev); yb = torch.empty(N, H, W, C, device=dev)
    be.bcast_fwd(v.data_ptr(), yb.data_ptr(), N, H * W, C, C)
    report('bcast fwd', yb, v.view(N, 1, 1, C).expand(N, H, W, C), 0)
    dv = torch.empty(N, C, device=dev)
    be.bcast_bwd(xh.data_ptr(), dv.data_ptr(), N, H * W, C, C)
    report('bcast bwd', dv, x.sum(dim=(2, 3)), 1e-5)
    # train BN fwd/bwd with relu + dropout mask
    C2 = 48
    xb = torch.randn(N, C2, H, W) * 2 + 0.5
    bn = torch.nn.BatchNorm2d(C2); bn.weight.data.uniform_(0.5, 1.5); bn.bias.data.normal_()
    bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 1.5)
    rm0, rv0 = bn.running_mean.clone(), bn.running_var.clone()
    drop = (torch.rand(N, H, W, C2) > 0.5).float()
    xr = xb.clone().requires_grad_(True)
    yref = torch.relu(bn(xr)) * drop.permute(0, 3, 1, 2) * 2.0
    dy = torch.randn(yref.shape); yref.backward(dy)
    rows = N * H * W
    xbd = nhwc(xb).to(dev)
    mean = torch.empty(C2, device=dev); rstd = torch.empty(C2, device=dev)
    rm, rv = rm0.to(dev), rv0.to(dev)
    g, b = bn.weight.data.to(dev), bn.bias.data.to(dev)
    be.bn_stats(xbd.data_ptr(), rows, C2, C2, 1e-5, 0.1, mean, rstd, rm, rv)
    yd = torch.empty(N, H, W, C2, device=dev); dropd = drop.to(dev)
    be.bn_apply(xbd.data_ptr(), rows, C2, C2, mean, rstd, g, b, True, dropd, 2.0, yd.data_ptr(), C2)
    report('bn train fwd', yd.permute(0, 3, 1, 2), yref.detach(), 2e-6)
    report('bn running_mean', rm, bn.running_mean, 1e-6); report('bn running_var', rv, bn.running_var, 1e-6)
    dx = torch.empty(N, H, W, C2, device=dev); dg = torch.zeros(C2, device=dev); db = torch.zeros(C2, device=dev)
    be.bn_bwd(nhwc(dy).to(dev).data_ptr(), C2, xbd.data_ptr(), C2, yd.data_ptr(), C2, rows, C2, mean, rstd, g, True, dropd, 2.0,
              dx.data_ptr(), C2, dg, db, False)
    report('bn bwd dx', dx.permute(0, 3, 1, 2), xr.grad, 1e-5)
    report('bn bwd dgamma', dg, bn.weight.grad, 1e-5); report('bn bwd dbeta', db, bn.bias.grad, 1e-5)
    # im2col stem
    xi = torch.randn(2, 3, 33, 29)
    xid = torch.empty(2, 33, 29, 4, device=dev)
    be.nchw_to_nhwc(xi.to(dev), xid.data_ptr(), 2, 3, 33, 29, 4)
    oh, ow = (33 + 6 - 7) // 2 + 1, (29 + 6 - 7) // 2 + 1
    col = torch.empty(2 * oh * ow, 160, device=dev)
    be.im2col(xid.data_ptr(), col.data_ptr(), 2, 33, 29, 3, 4, 7, 7, 2, 3, 1, oh, ow, 160)
    ref = F.unfold(xi, 7, dilation=1, padding=3, stride=2)       # (N, C*49, L) ordered c, r, s
    ref = ref.view(2, 3, 49, oh * ow).permute(0, 3, 2, 1).reshape(2 * oh * ow, 147)   # rows, (r s), c
    report('im2col', col[:, :147], ref, 0)
    print('im2col pad zero', float(col[:, 147:].abs().max()))
    # colsum, bn_fold
    out = torch.zeros(C, device=dev)
    be.colsum(xh.data_ptr(), C, N * H * W, C, out, False)
    report('colsum', out, x.sum(dim=(0, 2, 3)), 1e-5)


if __name__ == '__main__':
    print('SMs', L.call('b2_num_sms'), torch.cuda.get_device_name(0), flush=True)
    for g in sys.argv[1:]:
        print('==== group', g, flush=True)
        globals()['g_' + g]()

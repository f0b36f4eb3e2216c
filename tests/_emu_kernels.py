"""TEST DOUBLE — not part of the product.

Mirrors cutmix_semisup_seg_b200.kernels.ActKernels with torch CPU math so that the engine's graph
construction and hand-written backward tape (gradient accumulation, ReLU-gate / residual / BN-scale
folding, concat slices, strided dgrad phases) can be unit-tested in the GPU-less container.  Nothing in
the package imports this file; GPU tests exercise the real kernels.
"""
import torch
import torch.nn.functional as F

from cutmix_semisup_seg_b200.acts import Act

DT = torch.float64


def _v(a):
    """(N,C,H,W) float64 copy of an Act's logical contents."""
    return a.view4().permute(0, 3, 1, 2).to(DT)


def _store(a, t_nchw, accumulate=False):
    v = a.view4()
    val = t_nchw.permute(0, 2, 3, 1).to(torch.float32)
    if accumulate:
        v += val
    else:
        v.copy_(val)


def _w(w, cout, t, ldb, k):
    """weights from raw storage (cout, t, ldb) -> (cout, t, k)."""
    flat = torch.as_strided(w, (cout, t, ldb), (t * ldb, ldb, 1), storage_offset=w.storage_offset())
    return flat[..., :k].to(DT)


class EmuKernels(object):
    name = 'emu'

    def __init__(self):
        self.calls = []

    # ---- conv ---------------------------------------------------------------------------------
    def conv_fwd(self, x, w, cout, kh, kw, cin, ldb, stride, pad, dil, out, scale=None, shift=None, addend=None,
                 gate=None, relu=False):
        self.calls.append('conv_fwd')
        wt = _w(w, cout, kh * kw, ldb, cin).view(cout, kh, kw, cin).permute(0, 3, 1, 2)
        xin = _v(x)
        if x.h == 1 and x.n == 1 and kh == 1:        # flattened view
            pass
        y = F.conv2d(xin, wt, stride=stride, padding=pad, dilation=dil)
        if scale is not None:
            y = y * scale.to(DT).view(1, -1, 1, 1)
        if shift is not None:
            y = y + shift.to(DT).view(1, -1, 1, 1)
        if addend is not None:
            y = y + _v(addend)
        if relu:
            y = torch.relu(y)
        if gate is not None:
            y = y * (_v(gate) > 0)
        _store(out, y)

    def conv_dgrad(self, g, wt, cin, kh, kw, cout, ldb, stride, pad, dil, dx, addend=None, gate=None, accumulate=False,
                   want_stats=False, stats_sub=None):
        self.calls.append('conv_dgrad' + ('+stats' if want_stats else ''))
        w = _w(wt, cin, kh * kw, ldb, cout).view(cin, kh, kw, cout).permute(3, 0, 1, 2)   # (cout, cin, kh, kw)
        gy = _v(g)
        opad_h = dx.h - ((g.h - 1) * stride - 2 * pad + dil * (kh - 1) + 1)
        opad_w = dx.w - ((g.w - 1) * stride - 2 * pad + dil * (kw - 1) + 1)
        d = F.conv_transpose2d(gy, w, stride=stride, padding=pad, dilation=dil, output_padding=(opad_h, opad_w))
        if addend is not None:
            d = d + _v(addend)
        if gate is not None:
            d = d * (_v(gate) > 0)
        _store(dx, d, accumulate)
        if want_stats:
            assert gate is not None and not accumulate
            yv = _v(gate) if stats_sub is None else _v(gate) - _v(stats_sub)
            return (d.sum(dim=(0, 2, 3)), (d * yv).sum(dim=(0, 2, 3)))
        return None

    def stats_ok(self, t):
        return t.c % 4 == 0 and t.ld % 4 == 0 and t.off % 4 == 0

    def bn_eval_param_grad_from_stats(self, stats, gamma, beta, dgamma, dbeta, accumulate):
        self.calls.append('bn_eval_param_grad_from_stats')
        sg, sgy = stats
        dg = (sgy - beta.to(DT) * sg) / gamma.to(DT)
        for tgt, val in ((dgamma, dg), (dbeta, sg)):
            if accumulate:
                tgt += val.to(torch.float32)
            else:
                tgt.copy_(val.to(torch.float32))

    def bn_eval_param_grad_wdot(self, stats, g, w, gw, bn, dgamma, dbeta, accumulate):
        self.calls.append('bn_eval_param_grad_wdot' + ('+stats' if stats is not None else ''))
        sg = stats[0] if stats is not None else _v(g).sum(dim=(0, 2, 3))
        c = bn.weight.numel()
        db = (dbeta.to(DT) if accumulate else 0.0) + sg
        dot = (w.detach().to(DT).reshape(c, -1) * gw.to(DT).reshape(c, -1)).sum(dim=1)
        invstd = 1.0 / torch.sqrt(bn.running_var.to(DT) + bn.eps)
        dg = dot / bn.weight.detach().to(DT) - invstd * bn.running_mean.to(DT) * db
        dbeta.copy_(db.to(torch.float32)); dgamma.copy_(dg.to(torch.float32))

    def conv_wgrad(self, g, x, dw, cout, kh, kw, cin, stride, pad, dil, row_scale=None, accumulate=False):
        self.calls.append('conv_wgrad')
        xin = _v(x).requires_grad_(False)
        with torch.enable_grad():
            w = torch.zeros(cout, cin, kh, kw, dtype=DT, requires_grad=True)
            y = F.conv2d(xin, w, stride=stride, padding=pad, dilation=dil)
            y.backward(_v(g))
        gw = w.grad.permute(0, 2, 3, 1).reshape(cout, kh * kw, cin)
        if row_scale is not None:
            gw = gw * row_scale.to(DT).view(-1, 1, 1)
        flat = torch.as_strided(dw, (cout, kh * kw, cin), (kh * kw * cin, cin, 1), storage_offset=dw.storage_offset())
        if accumulate:
            flat += gw.to(torch.float32)
        else:
            flat.copy_(gw.to(torch.float32))

    def transpose_w(self, w, cout, t, cin, scale=None):
        ldb = (cout + 3) // 4 * 4
        src = _w(w, cout, t, cin, cin)
        if scale is not None:
            src = src * scale.to(DT).view(-1, 1, 1)
        out = torch.zeros(cin, t, ldb)
        out[..., :cout] = src.permute(2, 1, 0).to(torch.float32)
        return out, ldb

    # ---- other ops --------------------------------------------------------------------------------
    def nchw_to_act(self, x_nchw, ld):
        n, c, h, w = x_nchw.shape
        out = Act.alloc(n, h, w, c, x_nchw.device, ld=ld)
        out.base.zero_()
        out.view4().copy_(x_nchw.permute(0, 2, 3, 1))
        return out

    def im2col(self, x, kh, kw, stride, pad, dil, oh, ow, kpad):
        cols = F.unfold(_v(x), (kh, kw), dilation=dil, padding=pad, stride=stride)       # (N, C*kh*kw, L)
        n = x.n
        cols = cols.view(n, x.c, kh * kw, oh * ow).permute(0, 3, 2, 1).reshape(n * oh * ow, kh * kw * x.c)
        col = Act.alloc(1, 1, n * oh * ow, kpad, x.device)
        col.base.zero_()
        col.base.view(-1, kpad)[:, :kh * kw * x.c] = cols.to(torch.float32)
        return col

    # ---- U-Net decoder operators (csrc/unet.cu) ----------------------------------------------------
    def upsample2x_add(self, x, skip, out):
        self.calls.append('upsample2x_add')
        y = _v(x).repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
        if skip is not None:
            y = y + _v(skip)
        _store(out, y)

    def upsample2x_bwd(self, dy, dx, accumulate=False):
        self.calls.append('upsample2x_bwd')
        g = _v(dy)
        n, c, h2, w2 = g.shape
        _store(dx, g.view(n, c, h2 // 2, 2, w2 // 2, 2).sum(dim=(3, 5)), accumulate)

    def mul_mask(self, x, mask, scale, out):
        self.calls.append('mul_mask')
        _store(out, _v(x) * mask.permute(0, 3, 1, 2).to(DT) * scale)

    def avgpool2x2(self, x, out):
        self.calls.append('avgpool2x2')
        _store(out, F.avg_pool2d(_v(x), 2, 2))

    def avgpool2x2_bwd(self, dy, dx, accumulate=False):
        self.calls.append('avgpool2x2_bwd')
        g = _v(dy)
        full = torch.zeros(dx.n, dx.c, dx.h, dx.w, dtype=DT)
        up = g.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3) * 0.25
        full[:, :, :up.shape[2], :up.shape[3]] = up
        _store(dx, full, accumulate)

    def scale_channels(self, g, scale, dst, accumulate=False):
        self.calls.append('scale_channels')
        _store(dst, _v(g) * scale.to(DT).view(1, -1, 1, 1), accumulate)

    def col2im(self, dcol, dx, kh, kw, stride, pad, dil, oh, ow, kpad, accumulate=False):
        """b2_col2im's gather loop, written out tap by tap (independent of F.fold)."""
        self.calls.append('col2im')
        n, h, w, c = dx.n, dx.h, dx.w, dx.c
        cols = dcol.base.view(-1, kpad)[:, :kh * kw * c].to(DT).view(n, oh, ow, kh * kw, c)
        acc = torch.zeros(n, h, w, c, dtype=DT)
        for r in range(kh):
            for s in range(kw):
                for oy in range(oh):
                    y = oy * stride - pad + r * dil
                    if not 0 <= y < h:
                        continue
                    xs = torch.arange(ow) * stride - pad + s * dil
                    ok = (xs >= 0) & (xs < w)
                    acc[:, y, xs[ok]] += cols[:, oy, ok, r * kw + s]
        v = dx.view4()
        if accumulate:
            v += acc.to(torch.float32)
        else:
            v.copy_(acc.to(torch.float32))

    def act_to_nchw(self, a):
        return a.view4().permute(0, 3, 1, 2).contiguous().clone()

    def maxpool_fwd(self, x, out, idx):
        y, ind = F.max_pool2d(_v(x), 3, 2, 1, ceil_mode=False, return_indices=True) if False else (None, None)
        xv = _v(x)
        n, c, h, w = xv.shape
        padded = F.pad(xv, (1, 2, 1, 2), value=float('-inf'))
        best = torch.full((n, c, out.h, out.w), float('-inf'), dtype=DT)
        bi = torch.zeros((n, c, out.h, out.w), dtype=torch.uint8)
        for r in range(3):
            for s in range(3):
                cand = padded[:, :, r:r + 2 * out.h:2, s:s + 2 * out.w:2][:, :, :out.h, :out.w]
                upd = cand > best
                best = torch.where(upd, cand, best)
                bi = torch.where(upd, torch.tensor(r * 3 + s, dtype=torch.uint8), bi)
        _store(out, best)
        idx.copy_(bi.permute(0, 2, 3, 1))

    def maxpool_bwd(self, dy, idx, dx):
        g = _v(dy)
        n, c, oh, ow = g.shape
        d = torch.zeros(n, c, dx.h + 3, dx.w + 3, dtype=DT)
        ind = idx.permute(0, 3, 1, 2)
        for r in range(3):
            for s in range(3):
                sel = (ind == r * 3 + s).to(DT) * g
                d[:, :, r:r + 2 * oh:2, s:s + 2 * ow:2][:, :, :oh, :ow] += sel
        _store(dx, d[:, :, 1:1 + dx.h, 1:1 + dx.w])

    def bilinear_fwd(self, x, out, align_corners):
        _store(out, F.interpolate(_v(x), size=(out.h, out.w), mode='bilinear', align_corners=bool(align_corners)))

    def bilinear_fwd_nchw(self, x, out_nchw, align_corners):
        out_nchw.copy_(F.interpolate(_v(x), size=out_nchw.shape[2:4], mode='bilinear', align_corners=bool(align_corners)))

    def _bil_bwd(self, g_nchw, dx, align_corners, mul, accumulate):
        with torch.enable_grad():
            xz = torch.zeros(dx.n, dx.c, dx.h, dx.w, dtype=DT, requires_grad=True)
            y = F.interpolate(xz, size=g_nchw.shape[2:4], mode='bilinear', align_corners=bool(align_corners))
            y.backward(g_nchw.to(DT) * mul)
        _store(dx, xz.grad, accumulate)

    def bilinear_bwd(self, dy, dx, align_corners, accumulate=False):
        self._bil_bwd(_v(dy), dx, align_corners, 1.0, accumulate)

    def bilinear_bwd_nchw(self, dy_nchw, dx, align_corners, scale_dev=None, scale_host=1.0, accumulate=False):
        mul = (float(scale_dev[0]) if scale_dev is not None else 1.0) * scale_host
        self._bil_bwd(dy_nchw, dx, align_corners, mul, accumulate)

    def gap_fwd(self, x, out):
        _store(out, _v(x).mean(dim=(2, 3), keepdim=True))

    def gap_bwd(self, dy, dx, accumulate=False):
        _store(dx, (_v(dy) / (dx.h * dx.w)).expand(dx.n, dx.c, dx.h, dx.w), accumulate)

    def bcast_fwd(self, v, out):
        _store(out, _v(v).expand(out.n, out.c, out.h, out.w))

    def bcast_bwd(self, dy, dv):
        _store(dv, _v(dy).sum(dim=(2, 3), keepdim=True))

    def bn_stats(self, x, eps, momentum, mean, rstd, running_mean, running_var):
        xv = _v(x)
        m = xv.mean(dim=(0, 2, 3)); var = xv.var(dim=(0, 2, 3), unbiased=False)
        mean.copy_(m.to(torch.float32)); rstd.copy_((1.0 / torch.sqrt(var + eps)).to(torch.float32))
        n = x.rows
        if running_mean is not None:
            running_mean.mul_(1 - momentum).add_(momentum * m.to(torch.float32))
            unb = var * n / (n - 1) if n > 1 else var
            running_var.mul_(1 - momentum).add_(momentum * unb.to(torch.float32))

    def bn_apply(self, x, mean, rstd, gamma, beta, relu, dropmask, drop_scale, out, residual=None):
        sh = (1, -1, 1, 1)
        y = (_v(x) - mean.to(DT).view(sh)) * rstd.to(DT).view(sh) * gamma.to(DT).view(sh) + beta.to(DT).view(sh)
        if residual is not None:
            y = y + _v(residual)
        if relu:
            y = torch.relu(y)
        if dropmask is not None:
            y = y * dropmask.permute(0, 3, 1, 2).to(DT) * drop_scale
        _store(out, y)

    def bn_bwd(self, dy, x, y, mean, rstd, gamma, relu, dropmask, drop_scale, dx, dgamma, dbeta, accumulate_params, g_out=None,
               gate_beta=None):        # gate_beta: the device kernel may recompute the gate from x; the sign of y is the same
        sh = (1, -1, 1, 1)
        g = _v(dy)
        if relu:
            g = g * (_v(y) > 0)
        if dropmask is not None:
            g = g * dropmask.permute(0, 3, 1, 2).to(DT) * drop_scale
        if g_out is not None:
            _store(g_out, g)
        xhat = (_v(x) - mean.to(DT).view(sh)) * rstd.to(DT).view(sh)
        n = x.rows
        db = g.sum(dim=(0, 2, 3)); dg = (g * xhat).sum(dim=(0, 2, 3))
        d = gamma.to(DT).view(sh) * rstd.to(DT).view(sh) * (g - db.view(sh) / n - xhat * dg.view(sh) / n)
        _store(dx, d)
        for tgt, val in ((dgamma, dg), (dbeta, db)):
            if tgt is not None:
                if accumulate_params:
                    tgt += val.to(torch.float32)
                else:
                    tgt.copy_(val.to(torch.float32))

    def bn_fold(self, gamma, beta, mean, var, eps, scale, shift):
        s = gamma / torch.sqrt(var + eps)
        scale.copy_(s); shift.copy_(beta - mean * s)

    def bn_eval_param_grad(self, g, y, gamma, beta, sub, dgamma, dbeta, accumulate):
        self.calls.append('bn_eval_param_grad')
        gv = _v(g); yv = _v(y)
        if sub is not None:
            yv = yv - _v(sub)
        sg = gv.sum(dim=(0, 2, 3)); sgy = (gv * yv).sum(dim=(0, 2, 3))
        dg = (sgy - beta.to(DT) * sg) / gamma.to(DT)
        for tgt, val in ((dgamma, dg), (dbeta, sg)):
            if accumulate:
                tgt += val.to(torch.float32)
            else:
                tgt.copy_(val.to(torch.float32))

    def colsum(self, g, out, accumulate):
        v = _v(g).sum(dim=(0, 2, 3)).to(torch.float32)
        if accumulate:
            out += v
        else:
            out.copy_(v)

    def relu_gate(self, g, y):
        v = g.view4()
        v.mul_((y.view4() > 0).to(v.dtype))

    def copy_act(self, dst, src, accumulate=False):
        if accumulate:
            dst.view4().add_(src.view4())
        else:
            dst.view4().copy_(src.view4())

    def fill_act(self, a, value):
        a.view4().fill_(value)

    def dropout_mask(self, n, h, w, c, p, seed, offset, device, offset_dev=None):
        off = offset + (int(offset_dev[0]) if offset_dev is not None else 0)
        g = torch.Generator().manual_seed((seed + off) % (2 ** 31))
        return (torch.rand((n, h, w, c), generator=g) >= p).float()

    @staticmethod
    def empty(shape, device, dtype=torch.float32):
        return torch.empty(shape, device=device, dtype=dtype)

    def copy_rows(self, dst, ldd, src, lds, rows, c, accumulate=False):
        d = torch.as_strided(dst, (rows, c), (ldd, 1), storage_offset=dst.storage_offset())
        s_ = torch.as_strided(src, (rows, c), (lds, 1), storage_offset=src.storage_offset())
        if accumulate:
            d += s_
        else:
            d.copy_(s_)

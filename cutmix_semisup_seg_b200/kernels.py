"""Activation-level kernel interface used by the engine.

`ActKernels` maps operations on `Act` views (NHWC, leading dimension + channel offset) onto the
C-ABI launches of `ops.CudaBackend`.  It owns the launch-shape decisions that are not arithmetic:
flattening 1x1 stride-1 convolutions into plain GEMMs, decomposing strided input-gradients into
stride-1 phases, and the 3xTF32 precision mode (operand splitting).  There is no other
implementation in the product; tests/_emu_kernels.py mirrors this interface with torch CPU math
only to unit-test the engine's graph/backward logic without a GPU.
"""
import torch

from .acts import Act
from . import ops as O


def _pad4(v):
    return (v + 3) // 4 * 4


class ActKernels(object):
    name = 'cuda'
    KCHUNK = 1024           # parity mode (3xTF32): reduction terms per TMEM accumulation of a forward / dgrad GEMM

    def __init__(self, backend=None, n_split=1):
        self.be = backend if backend is not None else O.default_backend()
        self.n_split = n_split          # 1: single-pass TF32 (throughput); 3: 3xTF32 (parity mode)
        self.kchunk = 1024 if n_split > 1 else 0   # parity mode also bounds the pixels per TMEM accumulation in wgrad

    # ------------------------------------------------------------------------------ helpers
    def _split(self, ptr_tensor_like):
        raise NotImplementedError

    def _split_act(self, a):
        """Dense hi/lo copies of an Act (parity mode only)."""
        hi = Act.alloc(a.n, a.h, a.w, a.c, a.device, ld=_pad4(a.c))
        lo = Act.alloc(a.n, a.h, a.w, a.c, a.device, ld=_pad4(a.c))
        if a.ld == hi.ld and a.off == 0 and a.c == a.ld:
            h, l = self.be.split_tf32(a.base)
            return Act(h, a.n, a.h, a.w, a.c, a.ld), Act(l, a.n, a.h, a.w, a.c, a.ld)
        tmp = Act.alloc(a.n, a.h, a.w, a.c, a.device, ld=_pad4(a.c))
        if tmp.ld != a.c:
            self.be.fill(tmp.base, 0.0)
        self.be.slice_copy(tmp.ptr, tmp.ld, a.ptr, a.ld, a.rows, a.c)
        h, l = self.be.split_tf32(tmp.base)
        return Act(h, a.n, a.h, a.w, a.c, tmp.ld), Act(l, a.n, a.h, a.w, a.c, tmp.ld)

    def _split_w(self, w):
        return self.be.split_tf32(w)

    @staticmethod
    def _linear(*acts):
        """True if every Act can be addressed as one long row of pixels (always for NHWC + ld)."""
        return True

    # ------------------------------------------------------------------------------ convolution
    def _chunks(self, taps, k):
        """[(tap group, first channel, channels)] of one convolution GEMM.  Throughput mode: one launch.  Parity mode (3xTF32):
        tcgen05 accumulates in fp32 with TRUNCATION, so the error of one TMEM accumulation grows linearly with its reduction
        length (measured: 1.5e-6 of the output range at K = 64, 2e-5 at 2304, 1.4e-4 at 18432); reductions longer than KCHUNK
        terms therefore run as several launches over tap groups / channel ranges whose partial results are combined by the
        epilogue's round-to-nearest fp32 adds."""
        limit = self.KCHUNK if self.n_split > 1 else 0
        if not limit or k * len(taps) <= limit:
            return [(taps, 0, k)]
        if k > limit:
            return [([t], c0, min(limit, k - c0)) for t in taps for c0 in range(0, k, limit)]
        per = max(1, limit // k)
        return [(taps[i:i + per], 0, k) for i in range(0, len(taps), per)]

    def _gemm(self, a, a_lo, n, ih, iw, k, lda, b, b_lo, nb, tb, ldb, out_ptr, out_ld, oh, ow, taps, device, istride=1, **ep):
        """One logical convolution GEMM D = epilogue(sum_{tap, k} A B) through b2_conv_gemm (raw pointers; a_lo / b_lo: the
        low parts of the operand splits in parity mode).  Chunked launches use the linearity of the epilogue up to the ReLU:
        chunk 0 writes scale * acc_0 (+ addend) to a scratch buffer, the middle chunks accumulate scale * acc_i into it, the
        last launch adds it through the addend slot and applies shift / ReLU / gate / statistics / accumulate."""
        be = self.be
        base = dict(n_split=self.n_split)
        chunks = self._chunks(taps, k)
        if ep.get('want_stats'):
            ep = dict(ep, device=device)          # the statistics buffer is allocated by the (last) launch

        def ptrs(c_off):
            d = dict(base)
            d['a_lo_ptr'] = None if a_lo is None else a_lo + 4 * c_off
            d['b_lo_ptr'] = None if b_lo is None else b_lo + 4 * c_off
            return a + 4 * c_off, b + 4 * c_off, d
        if len(chunks) == 1:
            ap, bp, d = ptrs(0)
            return be.conv_gemm(ap, n, ih, iw, k, lda, bp, nb, tb, ldb, out_ptr, oh, ow, oh, ow, out_ld, taps, istride=istride,
                                **d, **ep)
        assert ep.get('scale2') is None, 'chunked launches do not support scale2'
        part_ld = _pad4(nb)
        part = torch.empty((n * oh * ow, part_ld), device=device, dtype=torch.float32)
        for ci, (grp, c_off, kc) in enumerate(chunks[:-1]):
            ap, bp, d = ptrs(c_off)
            d['scale'] = ep.get('scale')
            if ci == 0 and ep.get('addend') is not None:
                d['addend'] = ep['addend']; d['ld_add'] = ep['ld_add']
            be.conv_gemm(ap, n, ih, iw, kc, lda, bp, nb, tb, ldb, part.data_ptr(), oh, ow, oh, ow, part_ld, grp, istride=istride,
                         accumulate=ci > 0, **d)
        grp, c_off, kc = chunks[-1]
        ap, bp, d = ptrs(c_off)
        last = dict(ep)
        last['addend'] = part.data_ptr(); last['ld_add'] = part_ld
        return be.conv_gemm(ap, n, ih, iw, kc, lda, bp, nb, tb, ldb, out_ptr, oh, ow, oh, ow, out_ld, grp, istride=istride,
                            **d, **last)

    def conv_fwd(self, x, w, cout, kh, kw, cin, ldb, stride, pad, dil, out, scale=None, shift=None, addend=None,
                 gate=None, relu=False):
        """out = epilogue(conv(x, w)).  w: tensor whose storage is (cout, kh*kw, ldb)."""
        a_lo = b_lo = None
        xa, wb = x, w
        if self.n_split > 1:
            xa, xlo = self._split_act(x)
            wb, wlo = self._split_w(w)
            a_lo, b_lo = xlo.ptr, wlo.data_ptr()
        kwargs = dict(scale=scale, shift=shift, relu=relu)
        if addend is not None:
            kwargs['addend'] = addend.ptr; kwargs['ld_add'] = addend.ld
        if gate is not None:
            kwargs['gate'] = gate.ptr; kwargs['ld_gate'] = gate.ld
        taps = O.conv_taps(kh, kw, dil, pad)
        if kh == 1 and kw == 1 and stride == 1 and pad == 0:
            npix = x.rows
            self._gemm(xa.ptr, a_lo, 1, 1, npix, cin, xa.ld, wb.data_ptr(), b_lo, cout, 1, ldb, out.ptr, out.ld, 1, npix, taps,
                       x.device, **kwargs)
            return
        self._gemm(xa.ptr, a_lo, x.n, x.h, x.w, cin, xa.ld, wb.data_ptr(), b_lo, cout, kh * kw, ldb, out.ptr, out.ld, out.h, out.w,
                   taps, x.device, istride=stride, **kwargs)

    def conv_dgrad(self, g, wt, cin, kh, kw, cout, ldb, stride, pad, dil, dx, addend=None, gate=None, accumulate=False,
                   want_stats=False, stats_sub=None):
        """dx (+)= dgrad(g, W).  wt: transposed weights, storage (cin, kh*kw, ldb) with K = cout.
        Stride-1 convolutions fuse `addend` (partial gradient) and `gate` (ReLU of the producer);
        strided ones are decomposed into stride-1 phases scattered with output stride (no fusion).
        want_stats (needs gate): the epilogue also writes partial column sums of the gated gradient and of
        gradient * (gate - stats_sub); returned for bn_eval_param_grad_from_stats (frozen-BN parameter gradients of the
        layer that produced `gate` without another pass over HBM)."""
        be = self.be
        a_lo = b_lo = None
        ga, wb = g, wt
        if self.n_split > 1:
            ga, glo = self._split_act(g)
            wb, wlo = self._split_w(wt)
            a_lo, b_lo = glo.ptr, wlo.data_ptr()
        common = dict(a_lo_ptr=a_lo, b_lo_ptr=b_lo, n_split=self.n_split)
        if stride == 1:
            kwargs = dict(accumulate=accumulate)
            if addend is not None:
                kwargs['addend'] = addend.ptr; kwargs['ld_add'] = addend.ld
            if gate is not None:
                kwargs['gate'] = gate.ptr; kwargs['ld_gate'] = gate.ld
            if want_stats:
                assert gate is not None and not accumulate
                kwargs['want_stats'] = True
                if stats_sub is not None:
                    kwargs['stats_sub'] = stats_sub.ptr; kwargs['ld_stats_sub'] = stats_sub.ld
            taps = O.dgrad_taps(kh, kw, dil, pad)
            if kh == 1 and kw == 1 and pad == 0:
                npix = g.rows
                return self._gemm(ga.ptr, a_lo, 1, 1, npix, cout, ga.ld, wb.data_ptr(), b_lo, cin, 1, ldb, dx.ptr, dx.ld, 1, npix,
                                  taps, g.device, **kwargs)
            return self._gemm(ga.ptr, a_lo, g.n, g.h, g.w, cout, ga.ld, wb.data_ptr(), b_lo, cin, kh * kw, ldb, dx.ptr, dx.ld,
                              dx.h, dx.w, taps, g.device, **kwargs)
        assert addend is None and gate is None and not want_stats, 'strided dgrad is not fused'
        if not accumulate:
            self.fill_act(dx, 0.0)
        s = stride
        for a in range(s):
            for b in range(s):
                taps = []
                for r in range(kh):
                    if (a + pad - r * dil) % s:
                        continue
                    for q in range(kw):
                        if (b + pad - q * dil) % s:
                            continue
                        taps.append(((a + pad - r * dil) // s, (b + pad - q * dil) // s, r * kw + q))
                ph, pw = (dx.h - a + s - 1) // s, (dx.w - b + s - 1) // s
                if not taps or ph <= 0 or pw <= 0:
                    continue
                be.conv_gemm(ga.ptr, g.n, g.h, g.w, cout, ga.ld, wb.data_ptr(), cin, kh * kw, ldb, dx.ptr, ph, pw,
                             dx.h, dx.w, dx.ld, taps, ostride=s, ooh=a, oow=b, accumulate=True, **common)

    def conv_wgrad(self, g, x, dw, cout, kh, kw, cin, stride, pad, dil, row_scale=None, accumulate=False):
        """dw (+)= row_scale * wgrad(g, x).  dw: tensor whose storage is (cout, kh*kw, cin)."""
        be = self.be
        y_lo = x_lo = None
        ga, xa = g, x
        if self.n_split > 1:
            ga, glo = self._split_act(g)
            xa, xlo = self._split_act(x)
            y_lo, x_lo = glo.ptr, xlo.ptr
        taps = O.conv_taps(kh, kw, dil, pad)
        if kh == 1 and kw == 1 and stride == 1 and pad == 0:
            npix = x.rows
            be.conv_wgrad(ga.ptr, 1, 1, npix, cout, ga.ld, xa.ptr, 1, npix, cin, xa.ld, dw.data_ptr(), taps, 1,
                          accumulate=accumulate, dy_lo_ptr=y_lo, x_lo_ptr=x_lo, n_split=self.n_split, device=g.device,
                          row_scale=row_scale, kchunk=self.kchunk)
        else:
            be.conv_wgrad(ga.ptr, g.n, g.h, g.w, cout, ga.ld, xa.ptr, x.h, x.w, cin, xa.ld, dw.data_ptr(), taps, kh * kw,
                          istride=stride, accumulate=accumulate, dy_lo_ptr=y_lo, x_lo_ptr=x_lo, n_split=self.n_split,
                          device=g.device, row_scale=row_scale, kchunk=self.kchunk)

    # multi-tensor forms: one launch for all frozen BatchNorms of a network / all dgrad operands of a backward pass
    multi_tensor = True

    def bn_fold_multi(self, table, n_entries, max_c):
        self.be.bn_fold_multi(table, n_entries, max_c)

    def transpose_w_multi(self, table, n_entries, total_blocks):
        self.be.transpose_w_multi(table, n_entries, total_blocks)

    def transpose_w(self, w, cout, t, cin, scale=None):
        """(cout, t, cin) -> (cin, t, pad4(cout)) with optional per-cout scale.  Returns (tensor, ldb)."""
        ldb = _pad4(cout)
        return self.be.transpose_w(w, cout, t, cin, ldd=ldb, scale=scale), ldb

    # ------------------------------------------------------------------------------ other ops
    def nchw_to_act(self, x_nchw, ld):
        n, c, h, w = x_nchw.shape
        out = Act.alloc(n, h, w, c, x_nchw.device, ld=ld)
        self.be.nchw_to_nhwc(x_nchw.contiguous(), out.ptr, n, c, h, w, ld)
        return out

    def im2col(self, x, kh, kw, stride, pad, dil, oh, ow, kpad):
        col = Act.alloc(1, 1, x.n * oh * ow, kpad, x.device)
        self.be.im2col(x.ptr, col.ptr, x.n, x.h, x.w, x.c, x.ld, kh, kw, stride, pad, dil, oh, ow, kpad)
        return col

    def col2im(self, dcol, dx, kh, kw, stride, pad, dil, oh, ow, kpad, accumulate=False):
        """dx (+)= adjoint of im2col applied to dcol (input gradient of the stem; VAT only)."""
        self.be.col2im(dcol.ptr, dx.ptr, dx.n, dx.h, dx.w, dx.c, dx.ld, kh, kw, stride, pad, dil, oh, ow, kpad, accumulate)

    def act_to_nchw(self, a):
        out = torch.empty((a.n, a.c, a.h, a.w), device=a.device, dtype=torch.float32)
        self.be.nhwc_to_nchw(a.ptr, out, a.n, a.c, a.h, a.w, a.ld)
        return out

    # U-Net decoder operators (csrc/unet.cu)
    def upsample2x_add(self, x, skip, out):
        self.be.upsample2x_add(x.ptr, x.ld, None if skip is None else skip.ptr, 0 if skip is None else skip.ld, out.ptr, out.ld,
                               x.n, x.h, x.w, x.c)

    def upsample2x_bwd(self, dy, dx, accumulate=False):
        self.be.upsample2x_bwd(dy.ptr, dy.ld, dx.ptr, dx.ld, dx.n, dx.h, dx.w, dx.c, accumulate)

    def mul_mask(self, x, mask, scale, out):
        self.be.mul_mask(x.ptr, x.ld, mask, scale, out.ptr, out.ld, x.rows, x.c)

    def avgpool2x2(self, x, out):
        self.be.avgpool2x2(x.ptr, x.ld, out.ptr, out.ld, x.n, x.h, x.w, x.c)

    def avgpool2x2_bwd(self, dy, dx, accumulate=False):
        self.be.avgpool2x2_bwd(dy.ptr, dy.ld, dx.ptr, dx.ld, dx.n, dx.h, dx.w, dx.c, accumulate)

    def scale_channels(self, g, scale, dst, accumulate=False):
        self.be.scale_channels(g.ptr, g.ld, scale, dst.ptr, dst.ld, g.rows, g.c, accumulate)

    def maxpool_fwd(self, x, out, idx):
        assert x.ld == x.c and out.ld == out.c
        self.be.maxpool_fwd(x.ptr, out.ptr, idx.data_ptr(), x.n, x.h, x.w, x.c, out.h, out.w)

    def maxpool_bwd(self, dy, idx, dx):
        assert dy.ld == dy.c and dx.ld == dx.c
        self.be.maxpool_bwd(dy.ptr, idx.data_ptr(), dx.ptr, dx.n, dx.h, dx.w, dx.c, dy.h, dy.w)

    def bilinear_fwd(self, x, out, align_corners):
        self.be.bilinear_fwd(x.ptr, out.ptr, x.n, x.h, x.w, x.c, x.ld, out.h, out.w, out.ld, align_corners, False)

    def bilinear_fwd_nchw(self, x, out_nchw, align_corners):
        n, c, oh, ow = out_nchw.shape
        self.be.bilinear_fwd(x.ptr, out_nchw.data_ptr(), x.n, x.h, x.w, x.c, x.ld, oh, ow, 0, align_corners, True)

    def bilinear_bwd(self, dy, dx, align_corners, accumulate=False):
        self.be.bilinear_bwd(dy.ptr, dx.ptr, dx.n, dx.h, dx.w, dx.c, dx.ld, dy.h, dy.w, dy.ld, align_corners, False,
                             accumulate=accumulate)

    def bilinear_bwd_nchw(self, dy_nchw, dx, align_corners, scale_dev=None, scale_host=1.0, accumulate=False):
        n, c, oh, ow = dy_nchw.shape
        if max(oh, ow) <= 3072 and ow * 12 + dx.w * 40 <= 48 * 1024:      # table sizes of the separable kernels (netops.cu)
            self.be.bilinear_bwd_nchw(dy_nchw, dx.ptr, dx.n, dx.h, dx.w, dx.c, dx.ld, align_corners, scale_dev=scale_dev,
                                      scale_host=scale_host, accumulate=accumulate)
        else:
            self.be.bilinear_bwd(dy_nchw.data_ptr(), dx.ptr, dx.n, dx.h, dx.w, dx.c, dx.ld, oh, ow, 0, align_corners, True,
                                 scale_dev=scale_dev, scale_host=scale_host, accumulate=accumulate)

    # pooled / broadcast vectors are dense (N, C) in the C ABI
    def gap_fwd(self, x, out):
        assert out.ld == out.c, 'pooled vector must be dense'
        self.be.gap_fwd(x.ptr, out.ptr, x.n, x.h * x.w, x.c, x.ld)

    def gap_bwd(self, dy, dx, accumulate=False):
        assert dy.ld == dy.c, 'pooled vector must be dense'
        self.be.gap_bwd(dy.ptr, dx.ptr, dx.n, dx.h * dx.w, dx.c, dx.ld, accumulate=accumulate)

    def bcast_fwd(self, v, out):
        assert v.ld == v.c, 'broadcast vector must be dense'
        self.be.bcast_fwd(v.ptr, out.ptr, out.n, out.h * out.w, out.c, out.ld)

    def bcast_bwd(self, dy, dv):
        assert dv.ld == dv.c, 'broadcast vector must be dense'
        self.be.bcast_bwd(dy.ptr, dv.ptr, dy.n, dy.h * dy.w, dy.c, dy.ld)

    def bn_stats(self, x, eps, momentum, mean, rstd, running_mean, running_var):
        self.be.bn_stats(x.ptr, x.rows, x.c, x.ld, eps, momentum, mean, rstd, running_mean, running_var)

    def bn_apply(self, x, mean, rstd, gamma, beta, relu, dropmask, drop_scale, out, residual=None):
        self.be.bn_apply(x.ptr, x.rows, x.c, x.ld, mean, rstd, gamma, beta, relu, dropmask, drop_scale, out.ptr, out.ld,
                         res_ptr=None if residual is None else residual.ptr, ldr=0 if residual is None else residual.ld)

    def bn_bwd(self, dy, x, y, mean, rstd, gamma, relu, dropmask, drop_scale, dx, dgamma, dbeta, accumulate_params,
               g_out=None, gate_beta=None):
        self.be.bn_bwd(dy.ptr, dy.ld, x.ptr, x.ld, y.ptr, y.ld, x.rows, x.c, mean, rstd, gamma, relu, dropmask, drop_scale,
                       dx.ptr, dx.ld, dgamma, dbeta, accumulate_params,
                       g_out_ptr=None if g_out is None else g_out.ptr, ldgo=0 if g_out is None else g_out.ld,
                       gate_beta=gate_beta)

    def bn_fold(self, gamma, beta, mean, var, eps, scale, shift):
        self.be.bn_fold(gamma, beta, mean, var, eps, scale, shift)

    def bn_eval_param_grad(self, g, y, gamma, beta, sub, dgamma, dbeta, accumulate):
        self.be.bn_eval_param_grad(g.ptr, g.ld, y.ptr, y.ld, g.rows, g.c, gamma, beta, None, 0,
                                   None if sub is None else sub.ptr, 0 if sub is None else sub.ld, dgamma, dbeta,
                                   accumulate)

    def bn_eval_param_grad_from_stats(self, stats, gamma, beta, dgamma, dbeta, accumulate):
        self.be.bn_eval_param_grad_from_stats(stats, gamma, beta, dgamma, dbeta, accumulate)

    def bn_eval_param_grad_wdot(self, stats, g, w, gw, bn, dgamma, dbeta, accumulate):
        """dbeta (+)= sum_pix g; dgamma = <W, gw>/gamma - invstd*mean*dbeta (set from the accumulated totals).
        stats: what conv_dgrad(want_stats=True) returned for g, or None."""
        self.be.bn_eval_param_grad_wdot(stats, g.ptr, g.ld, g.rows, w, gw, bn.weight, bn.running_mean, bn.running_var,
                                        bn.eps, dgamma, dbeta, accumulate)

    def stats_ok(self, t):
        """Can a dgrad epilogue produce the column statistics of activation t's gradient?  (vector-path layout)"""
        return t.c % 4 == 0 and t.ld % 4 == 0 and t.off % 4 == 0

    def colsum(self, g, out, accumulate):
        self.be.colsum(g.ptr, g.ld, g.rows, g.c, out, accumulate)

    def relu_gate(self, g, y):
        self.be.relu_gate(g.ptr, g.ld, y.ptr, y.ld, g.rows, g.c)

    def copy_act(self, dst, src, accumulate=False):
        self.be.slice_copy(dst.ptr, dst.ld, src.ptr, src.ld, src.rows, src.c, accumulate)

    def copy_rows(self, dst, ldd, src, lds, rows, c, accumulate=False):
        """dst[r, :c] (+)= src[r, :c] on raw tensor storages with row pitches ldd / lds."""
        self.be.slice_copy(dst.data_ptr(), ldd, src.data_ptr(), lds, rows, c, accumulate)

    def fill_act(self, a, value):
        if a.ld == a.c and a.off == 0:
            self.be.fill(a.base, value)
        else:
            raise NotImplementedError('fill of a strided slice')

    def dropout_mask(self, n, h, w, c, p, seed, offset, device, offset_dev=None):
        mask = torch.empty((n, h, w, c), device=device, dtype=torch.float32)
        self.be.dropout_mask(mask, p, seed, offset, offset_dev=offset_dev)
        return mask

    # device-tensor helpers (allocation is plumbing, done through torch)
    @staticmethod
    def empty(shape, device, dtype=torch.float32):
        return torch.empty(shape, device=device, dtype=dtype)

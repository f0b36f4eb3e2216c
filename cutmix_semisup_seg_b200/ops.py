"""Python-side operator layer: thin wrappers that marshal torch CUDA tensors into the C ABI of
libb200seg.so (include/b200seg.h).  No arithmetic happens here; every op fails loudly if the
extension is missing or a tensor is not on a CUDA device (there is no CPU fallback).

Internal activation layout is NHWC fp32: a tensor of shape (N, H, W, ld) of which channels
[off, off + C) are addressed through views (`NHWC` below), so convolutions can write into slices
of a concat buffer.
"""
import ctypes

import numpy as np
import torch

from . import lib as L

LOSS_FNS = {'var': 0, 'logits_var': 1, 'logits_smoothl1': 2, 'bce': 3, 'kld': 4}


class CudaBackend(object):
    """The (only) product backend: every method is one or two launches of hand-written kernels."""

    name = 'cuda'

    def __init__(self):
        L.load()
        self._keep = []          # host-side arrays that must outlive an async launch (none today)
        self.launches = 0
        self._prof = None        # list of (kernel, flops, ev0, ev1) while profiling

    # ------------------------------------------------------------------ helpers
    def _s(self):
        return L.stream_ptr()

    def _call(self, name, *args):
        if self._prof is not None:
            return self._timed_call(name, 0.0, name, *args, _tag='')
        self.launches += 1
        return L.call(name, *args)

    # Profiling: CUDA events on the launching stream around every launch.  A device-side spin (~0.2 ms) is queued
    # ahead of the start event so that the host's enqueue latency (ctypes + tensor-map encode, ~50-100 us) is hidden
    # behind it: without the spin the GPU idles between `e0` and the launch and every short kernel reads ~0.1 ms.
    PROFILE_SPIN_CYCLES = 400000

    def _timed_call(self, kernel, flops, name, *args, _tag=None):
        """Launch with CUDA events recorded on the launching stream when profiling is on."""
        if self._prof is None:
            return self._call(name, *args)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(self.PROFILE_SPIN_CYCLES)
        e0.record()
        self.launches += 1
        rc = L.call(name, *args)
        e1.record()
        self._prof.append((kernel, flops, e0, e1, getattr(self, '_prof_tag', '') if _tag is None else _tag))
        return rc

    def start_profile(self):
        self._prof = []

    def stop_profile(self):
        """{kernel: {'ms': total device time, 'flops': algorithmic FLOPs, 'n': launches}}"""
        torch.cuda.synchronize()
        out, shapes = {}, {}
        for kernel, flops, e0, e1, tag in self._prof:
            ms = e0.elapsed_time(e1)
            d = out.setdefault(kernel, {'ms': 0.0, 'flops': 0.0, 'n': 0})
            d['ms'] += ms; d['flops'] += flops; d['n'] += 1
            d = shapes.setdefault(kernel + ' ' + tag, {'ms': 0.0, 'flops': 0.0, 'n': 0})
            d['ms'] += ms; d['flops'] += flops; d['n'] += 1
        self._prof = None
        self.last_shape_profile = shapes
        return out

    @staticmethod
    def empty(shape, device, dtype=torch.float32):
        return torch.empty(shape, device=device, dtype=dtype)

    @staticmethod
    def zeros(shape, device, dtype=torch.float32):
        return torch.zeros(shape, device=device, dtype=dtype)

    # ------------------------------------------------------------------ elementwise hot-path ops
    def ema_step_flat(self, tgt, src, alpha):
        L.require_cuda(tgt, src)
        L.require_dense(torch.float32, tgt, src)
        one_minus_alpha = 1.0 - alpha   # double, rounded to fp32 by ctypes like torch does (optim_weight_ema.py:22)
        self._call('b2_ema_step_flat', tgt.data_ptr(), src.data_ptr(), tgt.numel(), alpha, one_minus_alpha, self._s())

    def ema_step_table(self, table, n_chunks, alpha):
        one_minus_alpha = 1.0 - alpha
        self._call('b2_ema_step', table.data_ptr(), n_chunks, alpha, one_minus_alpha, self._s())

    def argmax_confusion(self, logits, labels, cm, ignore_value=255, want_pred=False):
        """cm (C*C int64, device) += confusion matrix of argmax(logits) vs labels; optionally returns the argmax map."""
        L.require_cuda(logits, labels, cm)
        n, c, h, w = logits.shape
        logits = logits.contiguous()
        L.require_dense(torch.float32, logits)
        L.require_dense(torch.int64, cm)
        if labels is not None:
            labels = labels.reshape(n, h, w).contiguous()
            assert labels.dtype == torch.int64
        pred = torch.empty((n, h, w), device=logits.device, dtype=torch.int64) if want_pred else None
        self._call('b2_argmax_confusion', logits.data_ptr(), L.ptr(labels), n, c, h * w,
                   -1 if ignore_value is None else int(ignore_value), L.ptr(cm), L.ptr(pred), self._s())
        return pred

    def box_mask_rasterize(self, boxes, h, w, init):
        """boxes: int32 CUDA tensor (N, B, 4) [y0,y1,x0,x1) -> (N,1,H,W) fp32."""
        L.require_cuda(boxes)
        L.require_dense(torch.int32, boxes)
        n, nb = boxes.shape[0], boxes.shape[1]
        out = torch.empty((n, 1, h, w), device=boxes.device, dtype=torch.float32)
        self._call('b2_box_mask_rasterize', boxes.data_ptr(), n, nb, h, w, float(init), out.data_ptr(), self._s())
        return out

    def mix(self, a, b, m, out=None):
        """out = a*(1-m) + b*m (b None: a*m).  a: (N,C,H,W) contiguous, m: (N,1,H,W)."""
        L.require_cuda(a, b, m)
        a = a.contiguous(); m = m.contiguous()
        if b is not None:
            b = b.contiguous()
        n, c, h, w = a.shape
        if tuple(m.shape) != (n, 1, h, w) or (b is not None and b.shape != a.shape):
            raise L.B2Error('mix: a/b must be (N,C,H,W) and the mask (N,1,H,W); got {} {} {}'.format(
                tuple(a.shape), None if b is None else tuple(b.shape), tuple(m.shape)))
        if out is None:
            out = torch.empty_like(a)
        L.require_dense(torch.float32, a, b, m, out)
        self._call('b2_mix', a.data_ptr(), L.ptr(b), m.data_ptr(), out.data_ptr(), n, c, h * w, self._s())
        return out

    def consistency(self, l0, l1, ls, m, lmask, loss_fn, conf_thresh, conf_per_pixel, ramp, cons_weight, dls=None):
        """Fused consistency loss.  Returns (out4, dls_unscaled): out4 = [loss, conf_rate, grad_scale, unsup_loss]."""
        L.require_cuda(l0, l1, ls, m, lmask)
        n, c, h, w = ls.shape
        hw = h * w
        if dls is None:
            dls = torch.empty_like(ls)
        L.require_dense(torch.float32, l0, l1, ls, m, lmask, dls)
        for t, shp in ((l0, ls.shape), (l1, ls.shape), (m, (n, 1, h, w)), (lmask, (n, 1, h, w)), (dls, ls.shape)):
            if t is not None and tuple(t.shape) != tuple(shp):
                raise L.B2Error('consistency: operand of shape {} where {} is expected'.format(tuple(t.shape), tuple(shp)))
        npart = L.call('b2_consistency_num_partials', n, hw)
        partials = torch.empty((npart * 3,), device=ls.device, dtype=torch.float64)
        out4 = torch.empty((4,), device=ls.device, dtype=torch.float32)
        self._call('b2_consistency_fwd_bwd', l0.data_ptr(), L.ptr(l1), ls.data_ptr(), L.ptr(m), L.ptr(lmask),
                   dls.data_ptr(), partials.data_ptr(), n, c, hw, LOSS_FNS[loss_fn], float(conf_thresh),
                   int(bool(conf_per_pixel)), self._s())
        self._call('b2_consistency_finalize', partials.data_ptr(), npart, n * hw, float(conf_thresh),
                   int(bool(conf_per_pixel)), float(ramp), float(cons_weight), out4.data_ptr(), self._s())
        return out4, dls

    def mix_per_sample(self, a, b, factors, out=None):
        """ICT image / valid-mask mix (train_seg_semisup_ict.py:310-311): out = a*(1-f) + b*f, f: (N,) fp32, one per sample."""
        L.require_cuda(a, b, factors)
        a = a.contiguous(); b = b.contiguous(); factors = factors.contiguous()
        n, c, h, w = a.shape
        assert factors.dtype == torch.float32 and factors.numel() == n
        if out is None:
            out = torch.empty_like(a)
        L.require_dense(torch.float32, a, b, factors, out)
        self._call('b2_mix_per_sample', a.data_ptr(), b.data_ptr(), factors.data_ptr(), out.data_ptr(), n, c, h * w, self._s())
        return out

    def ict_consistency(self, l0, l1, ls, factors, lmask, loss_fn, conf_thresh, conf_per_pixel, ramp, cons_weight, dls=None):
        """Fused ICT consistency block (train_seg_semisup_ict.py:318-392).  Returns (out4, dls_unscaled) like consistency()."""
        L.require_cuda(l0, l1, ls, factors, lmask)
        n, c, h, w = ls.shape
        hw = h * w
        factors = factors.contiguous()
        assert factors.dtype == torch.float32 and factors.numel() == n
        if dls is None:
            dls = torch.empty_like(ls)
        L.require_dense(torch.float32, l0, l1, ls, lmask, dls)
        confbar = None
        if conf_per_pixel and conf_thresh > 0.0:      # the reference's (N,N,1,H,W) broadcast: see include/b200seg.h
            confbar = torch.empty((hw,), device=ls.device, dtype=torch.float32)
            self._call('b2_ict_conf_mean', l0.data_ptr(), l1.data_ptr(), factors.data_ptr(), confbar.data_ptr(), n, c, hw,
                       float(conf_thresh), self._s())
        npart = L.call('b2_consistency_num_partials', n, hw)
        partials = torch.empty((npart * 3,), device=ls.device, dtype=torch.float64)
        out4 = torch.empty((4,), device=ls.device, dtype=torch.float32)
        self._call('b2_ict_consistency_fwd_bwd', l0.data_ptr(), l1.data_ptr(), ls.data_ptr(), factors.data_ptr(), L.ptr(lmask),
                   L.ptr(confbar), dls.data_ptr(), partials.data_ptr(), n, c, hw, LOSS_FNS[loss_fn], float(conf_thresh),
                   int(bool(conf_per_pixel)), self._s())
        self._call('b2_consistency_finalize', partials.data_ptr(), npart, n * hw, float(conf_thresh),
                   int(bool(conf_per_pixel)), float(ramp), float(cons_weight), out4.data_ptr(), self._s())
        return out4, dls

    def affine_grid_sample(self, x, theta, out_hw=None):
        """F.grid_sample(x, F.affine_grid(theta, ..., align_corners=True), align_corners=True), bilinear / zero padding
        (train_seg_semisup_aug_mt.py:302-306).  x: (N,C,H,W) fp32, theta: (N,2,3) fp32."""
        L.require_cuda(x, theta)
        x = x.contiguous(); theta = theta.contiguous()
        n, c, ih, iw = x.shape
        assert theta.dtype == torch.float32 and tuple(theta.shape) == (n, 2, 3)
        oh, ow = (ih, iw) if out_hw is None else out_hw
        y = torch.empty((n, c, oh, ow), device=x.device, dtype=torch.float32)
        L.require_dense(torch.float32, x, theta)
        self._call('b2_affine_grid_sample', x.data_ptr(), theta.data_ptr(), y.data_ptr(), n, c, ih, iw, oh, ow, self._s())
        return y

    def aug_consistency(self, ltea, ls, theta, um0, um1, loss_fn, conf_thresh, conf_per_pixel, ramp, cons_weight, dls=None):
        """Fused augmentation-consistency block (train_seg_semisup_aug_mt.py:291-391): teacher logits / probabilities / valid
        mask resampled into student space under the affine map `theta` inside the loss kernel.  Returns (out4, dls_unscaled)
        like consistency()."""
        L.require_cuda(ltea, ls, theta, um0, um1)
        n, c, h, w = ls.shape
        theta = theta.contiguous(); um0 = um0.contiguous(); um1 = um1.contiguous()
        assert theta.dtype == torch.float32 and tuple(theta.shape) == (n, 2, 3)
        assert tuple(ltea.shape) == tuple(ls.shape) and tuple(um0.shape) == (n, 1, h, w) and tuple(um1.shape) == (n, 1, h, w)
        if dls is None:
            dls = torch.empty_like(ls)
        L.require_dense(torch.float32, ltea, ls, theta, um0, um1, dls)
        npart = L.call('b2_consistency_num_partials', n, h * w)
        partials = torch.empty((npart * 3,), device=ls.device, dtype=torch.float64)
        out4 = torch.empty((4,), device=ls.device, dtype=torch.float32)
        self._call('b2_aug_consistency_fwd_bwd', ltea.data_ptr(), ls.data_ptr(), theta.data_ptr(), um0.data_ptr(),
                   um1.data_ptr(), dls.data_ptr(), partials.data_ptr(), n, c, h, w, LOSS_FNS[loss_fn], float(conf_thresh),
                   int(bool(conf_per_pixel)), self._s())
        self._call('b2_consistency_finalize', partials.data_ptr(), npart, n * h * w, float(conf_thresh),
                   int(bool(conf_per_pixel)), float(ramp), float(cons_weight), out4.data_ptr(), self._s())
        return out4, dls

    # ------------------------------------------------------------------ U-Net decoder (resunet.py / denseunet.py)
    def upsample2x_add(self, x_ptr, ldx, skip_ptr, lds, y_ptr, ldy, n, h, w, c):
        self._call('b2_upsample2x_add', x_ptr, ldx, skip_ptr, lds, y_ptr, ldy, n, h, w, c, self._s())

    def upsample2x_bwd(self, dy_ptr, lddy, dx_ptr, lddx, n, h, w, c, accumulate=False):
        self._call('b2_upsample2x_bwd', dy_ptr, lddy, dx_ptr, lddx, n, h, w, c, int(bool(accumulate)), self._s())

    def mul_mask(self, x_ptr, ldx, mask, scale, y_ptr, ldy, rows, c):
        self._call('b2_mul_mask', x_ptr, ldx, mask.data_ptr(), float(scale), y_ptr, ldy, rows, c, self._s())

    def avgpool2x2(self, x_ptr, ldx, y_ptr, ldy, n, ih, iw, c):
        self._call('b2_avgpool2x2', x_ptr, ldx, y_ptr, ldy, n, ih, iw, c, self._s())

    def avgpool2x2_bwd(self, dy_ptr, lddy, dx_ptr, lddx, n, ih, iw, c, accumulate=False):
        self._call('b2_avgpool2x2_bwd', dy_ptr, lddy, dx_ptr, lddx, n, ih, iw, c, int(bool(accumulate)), self._s())

    def scale_channels(self, g_ptr, ldg, scale, dst_ptr, ldd, rows, c, accumulate=False):
        self._call('b2_scale_channels', g_ptr, ldg, scale.data_ptr(), dst_ptr, ldd, rows, c, int(bool(accumulate)), self._s())

    # ------------------------------------------------------------------ data-format boundary (seg_transforms_cv.py:587-672)
    def normalize_to_tensor(self, img_u8, mean=None, std=None, out=None):
        """uint8 (N,H,W,3|4) pixels -> standardised fp32 (N,3,H,W) planes, bit-identical to the reference's numpy pipeline
        (img_as_float, (v - mean [* alpha]) / std in float64, .astype(float32)).  mean / std: sequences of 3 floats or None."""
        L.require_cuda(img_u8)
        assert img_u8.dtype == torch.uint8 and img_u8.dim() == 4
        img_u8 = img_u8.contiguous()
        n, h, w, cin = img_u8.shape
        if cin not in (3, 4):
            raise ValueError('image should have 3 channels, not {}'.format(cin))        # seg_transforms_cv.py:654
        if (mean is None) != (std is None):
            raise ValueError('mean and std must be given together')
        if out is None:
            out = torch.empty((n, 3, h, w), device=img_u8.device, dtype=torch.float32)
        m = s_ = None
        if mean is not None:
            m = (ctypes.c_double * 3)(*[float(v) for v in mean])
            s_ = (ctypes.c_double * 3)(*[float(v) for v in std])
        self._call('b2_normalize_to_tensor', img_u8.data_ptr(), n, h, w, cin, m, s_, out.data_ptr(), self._s())
        return out

    def labels_to_tensor(self, labels_u8):
        """uint8 (N,H,W) class indices (255 = ignore) -> int64 (N,1,H,W) (seg_transforms_cv.py:617)."""
        L.require_cuda(labels_u8)
        assert labels_u8.dtype == torch.uint8 and labels_u8.dim() == 3
        labels_u8 = labels_u8.contiguous()
        n, h, w = labels_u8.shape
        out = torch.empty((n, 1, h, w), device=labels_u8.device, dtype=torch.int64)
        self._call('b2_u8_to_tensor', labels_u8.data_ptr(), labels_u8.numel(), 0, out.data_ptr(), self._s())
        return out

    def mask_to_tensor(self, mask_u8):
        """uint8 (N,H,W) valid mask (0..255) -> fp32 (N,1,H,W) = float32(m * (1/255)) (seg_transforms_cv.py:620)."""
        L.require_cuda(mask_u8)
        assert mask_u8.dtype == torch.uint8 and mask_u8.dim() == 3
        mask_u8 = mask_u8.contiguous()
        n, h, w = mask_u8.shape
        out = torch.empty((n, 1, h, w), device=mask_u8.device, dtype=torch.float32)
        self._call('b2_u8_to_tensor', mask_u8.data_ptr(), mask_u8.numel(), 1, out.data_ptr(), self._s())
        return out

    def crop_flip_normalize(self, table, n, out_h, out_w, mean, std, want_labels, want_mask, device):
        """Fused pad / crop / flip / normalise-to-tensor gather (b2_crop_flip_normalize).  table: device uint8 tensor of n
        b2_crop_entry records.  Returns (image fp32 (n,3,h,w), labels int64 (n,1,h,w) | None, mask fp32 (n,1,h,w) | None)."""
        if (mean is None) != (std is None):
            raise ValueError('mean and std must be given together')
        image = torch.empty((n, 3, out_h, out_w), device=device, dtype=torch.float32)
        labels = torch.empty((n, 1, out_h, out_w), device=device, dtype=torch.int64) if want_labels else None
        mask = torch.empty((n, 1, out_h, out_w), device=device, dtype=torch.float32) if want_mask else None
        m = s_ = None
        if mean is not None:
            m = (ctypes.c_double * 3)(*[float(v) for v in mean])
            s_ = (ctypes.c_double * 3)(*[float(v) for v in std])
        self._call('b2_crop_flip_normalize', table.data_ptr(), int(n), int(out_h), int(out_w), m, s_, image.data_ptr(), L.ptr(labels),
                   L.ptr(mask), self._s())
        return image, labels, mask

    def crop_flip_u8(self, table, n, out_h, out_w, want_labels, want_mask, device):
        """b2_crop_flip_u8: the crop / flip gather with uint8 RGBA output (n, h, w, 4) for the colour-jitter branch."""
        image = torch.empty((n, out_h, out_w, 4), device=device, dtype=torch.uint8)
        labels = torch.empty((n, 1, out_h, out_w), device=device, dtype=torch.int64) if want_labels else None
        mask = torch.empty((n, 1, out_h, out_w), device=device, dtype=torch.float32) if want_mask else None
        self._call('b2_crop_flip_u8', table.data_ptr(), int(n), int(out_h), int(out_w), image.data_ptr(), L.ptr(labels), L.ptr(mask),
                   self._s())
        return image, labels, mask

    def geom_u8(self, entries, tables, n, out_h, out_w, want_labels, want_mask, device):
        """b2_geom_u8: scale / rotation crop (cv2.resize / cv2.warpAffine in OpenCV's fixed-point arithmetic) + flip of n
        variable-size uint8 samples -> RGBA uint8 (n, h, w, 4), labels int64 (n,1,h,w) | None, mask fp32 (n,1,h,w) | None.
        entries: device uint8 tensor of n b2_geom_entry records; tables: device int32 tensor."""
        assert entries.dtype == torch.uint8 and entries.numel() == n * 88 and tables.dtype == torch.int32
        assert tables.numel() >= n * 3 * (out_h + out_w)
        image = torch.empty((n, out_h, out_w, 4), device=device, dtype=torch.uint8)
        labels = torch.empty((n, 1, out_h, out_w), device=device, dtype=torch.int64) if want_labels else None
        mask = torch.empty((n, 1, out_h, out_w), device=device, dtype=torch.float32) if want_mask else None
        self._call('b2_geom_u8', entries.data_ptr(), tables.data_ptr(), int(n), int(out_h), int(out_w), image.data_ptr(), L.ptr(labels),
                   L.ptr(mask), self._s())
        return image, labels, mask

    def colour_jitter(self, img_u8, table_np):
        """In-place colour jitter of uint8 (N,H,W,3|4) pixels (b2_colour_jitter).  table_np: numpy structured array of N
        b2_colour_entry records (host); its device copy is made here."""
        L.require_cuda(img_u8)
        assert img_u8.dtype == torch.uint8 and img_u8.dim() == 4 and img_u8.shape[3] in (3, 4) and img_u8.is_contiguous()
        n, h, w, cs = img_u8.shape
        assert table_np.shape[0] == n and table_np.dtype.itemsize == 56
        host = np.ascontiguousarray(table_np)
        dev_tab = torch.from_numpy(host.view(np.uint8).copy()).to(img_u8.device)
        ws = torch.empty((n,), device=img_u8.device, dtype=torch.int64)
        self._call('b2_colour_jitter', img_u8.data_ptr(), n, h, w, cs, dev_tab.data_ptr(), host.ctypes.data, ws.data_ptr(), self._s())
        self.launches += 1
        return img_u8

    # ------------------------------------------------------------------ VAT (train_seg_semisup_vat_mt.py:214-301)
    def sample_l2norm(self, x):
        """mag[i] = sqrt(sum of squares of sample i) (normalize_eps, :217-219).  x: (N, ...) fp32 contiguous."""
        L.require_cuda(x)
        x = x.contiguous()
        L.require_dense(torch.float32, x)
        n = x.shape[0]
        per = x.numel() // n
        blocks = L.call('b2_sample_reduce_blocks', per)
        partials = torch.empty((n * blocks * 2,), device=x.device, dtype=torch.float64)
        mag = torch.empty((n,), device=x.device, dtype=torch.float32)
        self._call('b2_sample_l2norm', x.data_ptr(), n, per, partials.data_ptr(), mag.data_ptr(), self._s())
        return mag

    def vat_adaptive_radius(self, x, vat_radius):
        """radius[i] = vat_radius * sqrt(|vertical central differences|^2 + |horizontal ...|^2) * 0.5 (:289-296)."""
        L.require_cuda(x)
        x = x.contiguous()
        n, c, h, w = x.shape
        blocks = L.call('b2_sample_reduce_blocks', c * h * w)
        partials = torch.empty((n * blocks * 2,), device=x.device, dtype=torch.float64)
        radius = torch.empty((n,), device=x.device, dtype=torch.float32)
        self._call('b2_vat_adaptive_radius', x.data_ptr(), n, c, h, w, float(vat_radius), partials.data_ptr(),
                   radius.data_ptr(), self._s())
        return radius

    def add_scaled_per_sample(self, x, e, mag, radius, out=None):
        """out = x + (e / (mag[i] + 1e-12)) * radius  (radius: (N,) device tensor or a python float; x None: direction only)."""
        L.require_cuda(x, e, mag, radius if torch.is_tensor(radius) else None)
        e = e.contiguous()
        if x is not None:
            x = x.contiguous()
            assert x.shape == e.shape
        n = e.shape[0]
        per = e.numel() // n
        if out is None:
            out = torch.empty_like(e)
        r_dev = radius if torch.is_tensor(radius) else None
        L.require_dense(torch.float32, x, e, mag, r_dev, out)
        self._call('b2_add_scaled_per_sample', L.ptr(x), e.data_ptr(), mag.data_ptr(), L.ptr(r_dev),
                   0.0 if r_dev is not None else float(radius), out.data_ptr(), n, per, self._s())
        return out

    def col2im(self, dcol_ptr, dx_ptr, n, h, w, c, ldx, kh, kw, stride, pad, dil, oh, ow, kpad, accumulate=False):
        self._call('b2_col2im', dcol_ptr, dx_ptr, n, h, w, c, ldx, kh, kw, stride, pad, dil, oh, ow, kpad, int(bool(accumulate)),
                   self._s())

    def cross_entropy(self, logits, labels, ignore_index=255, dlogits=None):
        """Returns (out3, dlogits_unscaled): out3 = [loss, n_valid, grad_scale]."""
        L.require_cuda(logits, labels)
        n, c, h, w = logits.shape
        hw = h * w
        if dlogits is None:
            dlogits = torch.empty_like(logits)
        L.require_dense(torch.float32, logits, dlogits)
        L.require_dense(torch.int64, labels)
        if tuple(labels.shape) != (n, h, w):
            raise L.B2Error('cross_entropy: labels must be (N,H,W) int64, got {}'.format(tuple(labels.shape)))
        npart = L.call('b2_ce_num_partials', n, hw)
        partials = torch.empty((npart * 2,), device=logits.device, dtype=torch.float64)
        out3 = torch.empty((3,), device=logits.device, dtype=torch.float32)
        self._call('b2_ce_fwd_bwd', logits.data_ptr(), labels.data_ptr(), dlogits.data_ptr(), partials.data_ptr(),
                   n, c, hw, int(ignore_index), self._s())
        self._call('b2_ce_finalize', partials.data_ptr(), npart, out3.data_ptr(), self._s())
        return out3, dlogits

    def scale_inplace(self, x, scale_dev, scale_host=1.0):
        self._call('b2_scale_inplace', x.data_ptr(), x.numel(), L.ptr(scale_dev), float(scale_host), self._s())

    def add_inplace(self, dst, src):
        self._call('b2_add_inplace', dst.data_ptr(), src.data_ptr(), dst.numel(), self._s())

    def fill(self, dst, value):
        self._call('b2_fill', dst.data_ptr(), float(value), dst.numel(), self._s())

    def split_tf32(self, x):
        hi = torch.empty_like(x); lo = torch.empty_like(x)
        self._call('b2_split_tf32', x.data_ptr(), hi.data_ptr(), lo.data_ptr(), x.numel(), self._s())
        return hi, lo

    def transpose_w(self, w, a, t, b, out=None, ldd=None, scale=None):
        """(A,T,B) -> (B,T,ldd) on the raw storage of `w` (rows zero-padded to ldd, optional per-A scale)."""
        ldd = a if ldd is None else ldd
        if out is None:
            out = torch.empty((b, t, ldd), device=w.device, dtype=torch.float32)
        self._call('b2_transpose_w', w.data_ptr(), out.data_ptr(), a, t, b, ldd, L.ptr(scale), self._s())
        return out

    # ------------------------------------------------------------------ multi-tensor derived-weight kernels (csrc/multi.cu)
    def bn_fold_multi(self, table, n_entries, max_c):
        """table: device uint8 tensor holding n_entries b2_bn_fold_entry records (include/b200seg.h)."""
        self._call('b2_bn_fold_multi', table.data_ptr(), int(n_entries), int(max_c), self._s())

    def transpose_w_multi(self, table, n_entries, total_blocks):
        """table: device uint8 tensor holding n_entries b2_transpose_entry records."""
        self._call('b2_transpose_w_multi', table.data_ptr(), int(n_entries), int(total_blocks), self._s())

    # ------------------------------------------------------------------ tensor-core convolution
    def conv_gemm(self, a_ptr, n, ih, iw, k, lda, b_ptr, nb, tb, ldb, d_ptr, oh, ow, fh, fw, ldd, taps,
                  ostride=1, ooh=0, oow=0, istride=1, scale=None, shift=None, addend=None, ld_add=0,
                  gate=None, ld_gate=0, scale2=None, relu=False, accumulate=False,
                  a_lo_ptr=None, b_lo_ptr=None, n_split=1, max_ctas=0, want_stats=False, stats_sub=None,
                  ld_stats_sub=0, device=None):
        """Raw-pointer form of b2_conv_gemm (see include/b200seg.h).  want_stats: also return the fused column
        statistics buffer (stats tensor (row blocks, 2, ld), row blocks, ld) for b2_bn_eval_param_grad_from_stats."""
        taps_arr = np.ascontiguousarray(np.asarray(taps, dtype=np.int32).reshape(-1, 3))
        p = L.ConvParams()
        p.a = a_ptr; p.a_lo = a_lo_ptr; p.b = b_ptr; p.b_lo = b_lo_ptr; p.d = d_ptr
        p.n, p.ih, p.iw, p.k, p.lda = n, ih, iw, k, lda
        p.nb, p.tb, p.ldb = nb, tb, ldb
        p.oh, p.ow = oh, ow
        p.fh, p.fw, p.ldd, p.ostride, p.ooh, p.oow = fh, fw, ldd, ostride, ooh, oow
        p.istride = istride
        p.n_taps = taps_arr.shape[0]
        p.taps = taps_arr.ctypes.data
        p.scale = L.ptr(scale); p.shift = L.ptr(shift)
        p.addend = addend; p.ld_add = ld_add
        p.gate = gate; p.ld_gate = ld_gate
        p.scale2 = L.ptr(scale2)
        p.relu = int(bool(relu)); p.accumulate = int(bool(accumulate)); p.n_split = n_split
        p.max_ctas = max_ctas
        stats = None
        if want_stats:
            rows = int(L.call('b2_conv_stats_rows', ctypes.byref(p)))
            ld = (nb + 3) // 4 * 4
            buf = torch.empty((rows, 2, ld), device=device, dtype=torch.float32)
            p.stats = buf.data_ptr(); p.ld_stats = ld
            p.stats_sub = stats_sub; p.ld_stats_sub = ld_stats_sub
            stats = (buf, rows, ld)
        flops = 2.0 * n * oh * ow * nb * k * taps_arr.shape[0] * n_split
        self._prof_tag = 'pix{} k{} n{} taps{} s{}'.format(n * oh * ow, k, nb, taps_arr.shape[0], istride)
        self._timed_call('conv_gemm2_kernel' if (nb > 224 and max_ctas != 1) else 'conv_gemm_kernel', flops / n_split, 'b2_conv_gemm', ctypes.byref(p), self._s())
        return stats

    def conv_wgrad(self, dy_ptr, n, oh, ow, m, ldy, x_ptr, ih, iw, c, ldx, dw_ptr, taps, tw, istride=1,
                   accumulate=False, dy_lo_ptr=None, x_lo_ptr=None, n_split=1, max_ctas=0, device=None,
                   row_scale=None, kchunk=0):
        taps_arr = np.ascontiguousarray(np.asarray(taps, dtype=np.int32).reshape(-1, 3))
        p = L.WgradParams()
        p.dy = dy_ptr; p.dy_lo = dy_lo_ptr; p.x = x_ptr; p.x_lo = x_lo_ptr; p.dw = dw_ptr
        p.n, p.oh, p.ow, p.m, p.ldy = n, oh, ow, m, ldy
        p.ih, p.iw, p.c, p.ldx = ih, iw, c, ldx
        p.istride = istride
        p.n_taps = taps_arr.shape[0]
        p.taps = taps_arr.ctypes.data
        p.tw = tw
        p.accumulate = int(bool(accumulate)); p.n_split = n_split
        p.max_ctas = max_ctas
        p.row_scale = L.ptr(row_scale)
        p.kchunk = kchunk
        need = L.call('b2_conv_wgrad_workspace', ctypes.byref(p))
        ws = None
        if need > 0:
            ws = self._workspace(need, device)
            p.workspace = ws.data_ptr(); p.workspace_bytes = ws.numel()
        flops = 2.0 * n * oh * ow * m * c * taps_arr.shape[0]
        self._prof_tag = 'pix{} m{} c{} taps{} s{}'.format(n * oh * ow, m, c, taps_arr.shape[0], istride)
        self._timed_call('conv_wgrad_kernel', flops, 'b2_conv_wgrad', ctypes.byref(p), self._s())
        self.launches += 1 if need > 0 else 0

    # ------------------------------------------------------------------ NHWC network ops (raw pointers)
    def nchw_to_nhwc(self, src, dst_ptr, n, c, h, w, ldd):
        self._call('b2_nchw_to_nhwc', src.data_ptr(), dst_ptr, n, c, h, w, ldd, self._s())

    def nhwc_to_nchw(self, src_ptr, dst, n, c, h, w, lds):
        self._call('b2_nhwc_to_nchw', src_ptr, dst.data_ptr(), n, c, h, w, lds, self._s())

    def im2col(self, x_ptr, col_ptr, n, h, w, c, ldx, kh, kw, stride, pad, dil, oh, ow, kpad):
        self._call('b2_im2col', x_ptr, col_ptr, n, h, w, c, ldx, kh, kw, stride, pad, dil, oh, ow, kpad, self._s())

    def maxpool_fwd(self, x_ptr, y_ptr, idx_ptr, n, h, w, c, oh, ow):
        self._call('b2_maxpool3x3s2', x_ptr, y_ptr, idx_ptr, n, h, w, c, oh, ow, self._s())

    def maxpool_bwd(self, dy_ptr, idx_ptr, dx_ptr, n, h, w, c, oh, ow):
        self._call('b2_maxpool3x3s2_bwd', dy_ptr, idx_ptr, dx_ptr, n, h, w, c, oh, ow, self._s())

    def bilinear_fwd(self, x_ptr, y_ptr, n, ih, iw, c, ldx, oh, ow, ldy, align_corners, to_nchw):
        self._call('b2_bilinear_fwd', x_ptr, y_ptr, n, ih, iw, c, ldx, oh, ow, ldy, int(align_corners), int(to_nchw), self._s())

    def bilinear_bwd(self, dy_ptr, dx_ptr, n, ih, iw, c, ldx, oh, ow, ldy, align_corners, from_nchw, scale_dev=None,
                     scale_host=1.0, accumulate=False):
        self._call('b2_bilinear_bwd', dy_ptr, dx_ptr, n, ih, iw, c, ldx, oh, ow, ldy, int(align_corners), int(from_nchw),
                   L.ptr(scale_dev), float(scale_host), int(accumulate), self._s())

    def bilinear_bwd_nchw(self, dy, dx_ptr, n, ih, iw, c, ldx, align_corners, scale_dev=None, scale_host=1.0, accumulate=False):
        """Backward of the final resize: dy (N,C,OH,OW) contiguous -> NHWC dx, separable two-pass kernel."""
        oh, ow = dy.shape[2], dy.shape[3]
        need = int(L.call('b2_bilinear_bwd_nchw_workspace_floats', n, c, iw, oh))
        tmp = torch.empty((need,), device=dy.device, dtype=torch.float32)
        self._call('b2_bilinear_bwd_nchw', dy.data_ptr(), dx_ptr, tmp.data_ptr(), n, ih, iw, c, ldx, oh, ow, int(align_corners),
                   L.ptr(scale_dev), float(scale_host), int(accumulate), self._s())
        self.launches += 1

    def gap_fwd(self, x_ptr, y_ptr, n, hw, c, ldx):
        self._call('b2_gap_fwd', x_ptr, y_ptr, n, hw, c, ldx, self._s())

    def gap_bwd(self, dy_ptr, dx_ptr, n, hw, c, ldx, accumulate=False):
        self._call('b2_gap_bwd', dy_ptr, dx_ptr, n, hw, c, ldx, int(accumulate), self._s())

    def bcast_fwd(self, v_ptr, y_ptr, n, hw, c, ldy):
        self._call('b2_bcast_fwd', v_ptr, y_ptr, n, hw, c, ldy, self._s())

    def bcast_bwd(self, dy_ptr, dv_ptr, n, hw, c, ldy):
        self._call('b2_bcast_bwd', dy_ptr, dv_ptr, n, hw, c, ldy, self._s())

    def _red_ws(self, rows, c, device):
        need = L.call('b2_bn_workspace_doubles', rows, c)
        if getattr(self, '_rws', None) is None or self._rws.numel() < need or self._rws.device != torch.device(device):
            self._rws = torch.empty((int(need),), device=device, dtype=torch.float64)
        return self._rws

    def bn_stats(self, x_ptr, rows, c, ldx, eps, momentum, mean, rstd, running_mean, running_var):
        ws = self._red_ws(rows, c, mean.device)
        self._call('b2_bn_stats', x_ptr, rows, c, ldx, float(eps), float(momentum), mean.data_ptr(), rstd.data_ptr(),
                   L.ptr(running_mean), L.ptr(running_var), ws.data_ptr(), self._s())
        self.launches += 2

    def bn_apply(self, x_ptr, rows, c, ldx, mean, rstd, gamma, beta, relu, dropmask, drop_scale, y_ptr, ldy,
                 res_ptr=None, ldr=0):
        self._call('b2_bn_apply', x_ptr, rows, c, ldx, mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                   int(bool(relu)), L.ptr(dropmask), float(drop_scale), y_ptr, ldy, res_ptr, ldr, self._s())

    def bn_bwd(self, dy_ptr, lddy, x_ptr, ldx, y_ptr, ldy, rows, c, mean, rstd, gamma, relu, dropmask, drop_scale,
               dx_ptr, lddx, dgamma, dbeta, accumulate_params, g_out_ptr=None, ldgo=0, gate_beta=None):
        """gate_beta: the layer's beta when the ReLU gate may be recomputed from x (no residual): y is then not read."""
        ws = self._red_ws(rows, c, mean.device)
        self._call('b2_bn_bwd', dy_ptr, lddy, x_ptr, ldx, y_ptr, ldy, rows, c, mean.data_ptr(), rstd.data_ptr(),
                   gamma.data_ptr(), int(bool(relu)), L.ptr(dropmask), float(drop_scale), dx_ptr, lddx,
                   L.ptr(dgamma), L.ptr(dbeta), int(bool(accumulate_params)), g_out_ptr, ldgo, L.ptr(gate_beta),
                   ws.data_ptr(), self._s())
        self.launches += 3

    def bn_fold(self, gamma, beta, mean, var, eps, scale, shift):
        self._call('b2_bn_fold', gamma.data_ptr(), beta.data_ptr(), mean.data_ptr(), var.data_ptr(), float(eps),
                   scale.data_ptr(), shift.data_ptr(), gamma.numel(), self._s())

    def bn_eval_param_grad(self, dy_ptr, lddy, ybn_ptr, ldy, rows, c, gamma, beta, gate_ptr, ldg, sub_ptr, lds,
                           dgamma, dbeta, accumulate):
        ws = self._red_ws(rows, c, gamma.device)
        self._call('b2_bn_eval_param_grad', dy_ptr, lddy, ybn_ptr, ldy, rows, c, gamma.data_ptr(), beta.data_ptr(),
                   gate_ptr, ldg, sub_ptr, lds, dgamma.data_ptr(), dbeta.data_ptr(), int(bool(accumulate)),
                   ws.data_ptr(), self._s())
        self.launches += 2

    def bn_eval_param_grad_from_stats(self, stats, gamma, beta, dgamma, dbeta, accumulate):
        buf, rows, ld = stats
        c = gamma.numel()
        need = int(L.call('b2_bn_stats_workspace_doubles', c))
        if getattr(self, '_rws', None) is None or self._rws.numel() < need or self._rws.device != gamma.device:
            self._rws = torch.empty((need,), device=gamma.device, dtype=torch.float64)
        self._call('b2_bn_eval_param_grad_from_stats', buf.data_ptr(), rows, ld, c, gamma.data_ptr(), beta.data_ptr(),
                   dgamma.data_ptr(), dbeta.data_ptr(), int(bool(accumulate)), self._rws.data_ptr(), self._s())
        self.launches += 1

    def bn_eval_param_grad_wdot(self, stats, dy_ptr, lddy, rows, w, gw, gamma, mean, var, eps, dgamma, dbeta, accumulate):
        """Frozen-BN dgamma / dbeta from <W, dW> (b2_bn_eval_param_grad_wdot*).  stats: fused epilogue partial sums
        (buffer, row blocks, ld) or None -> sum_pix g is reduced from dy."""
        c = gamma.numel()
        row_len = w.numel() // c
        if stats is not None:
            buf, srows, ld = stats
            need = int(L.call('b2_bn_stats_workspace_doubles', c))
            if getattr(self, '_rws', None) is None or self._rws.numel() < need or self._rws.device != gamma.device:
                self._rws = torch.empty((need,), device=gamma.device, dtype=torch.float64)
            self._call('b2_bn_eval_param_grad_wdot_from_stats', buf.data_ptr(), srows, ld, c, w.data_ptr(), gw.data_ptr(),
                       row_len, gamma.data_ptr(), mean.data_ptr(), var.data_ptr(), float(eps), dgamma.data_ptr(),
                       dbeta.data_ptr(), int(bool(accumulate)), self._rws.data_ptr(), self._s())
            self.launches += 1
        else:
            ws = self._red_ws(rows, c, gamma.device)
            self._call('b2_bn_eval_param_grad_wdot', dy_ptr, lddy, rows, c, w.data_ptr(), gw.data_ptr(), row_len,
                       gamma.data_ptr(), mean.data_ptr(), var.data_ptr(), float(eps), dgamma.data_ptr(), dbeta.data_ptr(),
                       int(bool(accumulate)), ws.data_ptr(), self._s())
            self.launches += 2

    def colsum(self, dy_ptr, ld, rows, c, out, accumulate):
        ws = self._red_ws(rows, c, out.device)
        self._call('b2_colsum', dy_ptr, ld, rows, c, out.data_ptr(), int(bool(accumulate)), ws.data_ptr(), self._s())
        self.launches += 2

    def relu_gate(self, g_ptr, ldg, y_ptr, ldy, rows, c):
        self._call('b2_relu_gate', g_ptr, ldg, y_ptr, ldy, rows, c, self._s())

    def slice_copy(self, dst_ptr, ldd, src_ptr, lds, rows, c, accumulate=False):
        self._call('b2_slice_copy', dst_ptr, ldd, src_ptr, lds, rows, c, int(bool(accumulate)), self._s())

    def dropout_mask(self, mask, p, seed, offset, offset_dev=None):
        self._call('b2_dropout_mask', mask.data_ptr(), mask.numel(), float(p), int(seed), int(offset), L.ptr(offset_dev),
                   self._s())

    _ws = None

    def _workspace(self, nbytes, device):
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != torch.device(device):
            self._ws = torch.empty((int(nbytes),), device=device, dtype=torch.uint8)
        return self._ws


def conv_taps(kh, kw, dil, pad):
    """Tap table (dh, dw, weight_tap) of a forward convolution (input offset of each filter tap)."""
    return [(r * dil - pad, s * dil - pad, r * kw + s) for r in range(kh) for s in range(kw)]


def dgrad_taps(kh, kw, dil, pad):
    """Tap table of the stride-1 input-gradient convolution: dX[h] += dY[h + pad - r*dil] * W[r]."""
    return [(pad - r * dil, pad - s * dil, r * kw + s) for r in range(kh) for s in range(kw)]


_default = None


def default_backend():
    """Process-wide CudaBackend (raises if libb200seg.so is missing: no CPU fallback)."""
    global _default
    if _default is None:
        _default = CudaBackend()
    return _default


def set_default_backend(backend):
    """Test hook (like netbase.set_kernels_factory): tests/_emu_backend.py swaps in a torch-CPU double of this op
    INTERFACE so that the iteration's host logic (step.py, gradient all-reduce) runs without a GPU.  Returns the previous
    backend object (None: the CUDA backend has not been created yet); the product never calls this."""
    global _default
    prev, _default = _default, backend
    return prev

"""torch-CPU double of the step-level op INTERFACE (cutmix_semisup_seg_b200.ops.CudaBackend: mix, losses, masks, fills) for
host-logic tests of the training iteration without a GPU.  It follows the KERNELS' algorithms (un-scaled logit gradients +
device scalars [loss, conf_rate, grad_scale, unsup_loss]; ICT's batch-mean confidence map), not the oracle's autograd
formulation, so tests/test_step_emu.py compares two independent statements of the reference's loss blocks.
Never imported by the product."""
import math

import numpy as np
import torch
import torch.nn.functional as F


def _mix1(a, b, m):
    return a * (1.0 - m) + b * m              # fp32 tensors: one rounding per operation, like mix1() in elementwise.cu


class EmuBackend(object):
    name = 'emu-cpu'

    def __init__(self):
        self.launches = 0

    # ------------------------------------------------------------------ elementwise
    def fill(self, dst, value):
        dst.fill_(value); self.launches += 1

    def mix(self, a, b, m, out=None):
        r = a * m if b is None else _mix1(a, b, m)
        self.launches += 1
        if out is not None:
            out.copy_(r); return out
        return r

    def mix_per_sample(self, a, b, factors, out=None):
        r = _mix1(a, b, factors.reshape(-1, 1, 1, 1))
        self.launches += 1
        if out is not None:
            out.copy_(r); return out
        return r

    def box_mask_rasterize(self, boxes, h, w, init):
        n, nb, _ = boxes.shape
        m = torch.full((n, 1, h, w), float(init))
        for i in range(n):
            for j in range(nb):
                y0, y1, x0, x1 = (int(v) for v in boxes[i, j])
                m[i, 0, y0:y1, x0:x1] = 1.0 - m[i, 0, y0:y1, x0:x1]          # XOR toggle (mask_gen.py:114-116)
        self.launches += 1
        return m

    # ------------------------------------------------------------------ losses
    def cross_entropy(self, logits, labels, ignore_index=255, dlogits=None):
        n, c, h, w = logits.shape
        valid = labels != ignore_index
        logp = F.log_softmax(logits, dim=1)
        lab = labels.clamp(0, c - 1)
        nll = -logp.gather(1, lab[:, None])[:, 0]
        b = valid.sum().double()
        loss = (nll.double() * valid).sum() / b
        g = logp.exp()
        g = g - F.one_hot(lab, c).permute(0, 3, 1, 2).to(g.dtype)
        g = g * valid[:, None]
        self.launches += 2
        out3 = torch.tensor([float(loss), float(b), float(1.0 / b) if b > 0 else 0.0])
        return out3, g.contiguous()

    @staticmethod
    def _q_and_grad(fn, pt, lt, st, c):
        """Per-pixel loss q (N,1,H,W) and dq/d(student logits) for teacher probabilities pt / teacher logits lt."""
        ps = F.softmax(st, dim=1)
        if fn == 'var':
            d = ps - pt
            q = (d * d).sum(1, keepdim=True)
            g = 2.0 * d
            g = ps * (g - (g * ps).sum(1, keepdim=True))
        elif fn == 'logits_var':
            inv = 1.0 / math.sqrt(c)
            d = st - lt
            q = (d * d).sum(1, keepdim=True) * inv
            g = 2.0 * d * inv
        elif fn == 'logits_smoothl1':
            inv = 1.0 / math.sqrt(c)
            d = st - lt
            ad = d.abs()
            q = torch.where(ad < 1.0, 0.5 * d * d, ad - 0.5).sum(1, keepdim=True) * inv
            g = torch.where(ad < 1.0, d, torch.sign(d)) * inv
        elif fn == 'bce':
            eps = 1e-6
            inv_t, inv_p = 1.0 - pt, 1.0 - ps + eps
            q = (-(pt * torch.log(ps + eps) + inv_t * torch.log(inv_p))).sum(1, keepdim=True)
            g = -(pt / (ps + eps) - inv_t / inv_p)
            g = ps * (g - (g * ps).sum(1, keepdim=True))
        elif fn == 'kld':
            logp = F.log_softmax(st, dim=1)
            tlogt = torch.where(pt > 0, pt * torch.log(pt.clamp_min(1e-45)), torch.zeros_like(pt))
            q = (tlogt - pt * logp).sum(1, keepdim=True)
            g = ps * pt.sum(1, keepdim=True) - pt
        else:
            raise ValueError(fn)
        return q, g

    def _finish(self, q, g, w, conf, cw, conf_thresh, conf_per_pixel, ramp, cons_weight):
        """consistency_kernel's tail + consistency_finalize_kernel: conf = this pixel's mask (conf_rate), cw = the weight
        the per-pixel mode applies (conf itself, or ICT's batch-mean map)."""
        P = float(q.numel())
        a = conf.double().sum()
        b = (q.double() * w.double()).sum()
        cc = (q.double() * w.double() * cw.double()).sum()
        conf_rate = a / P
        if conf_thresh > 0.0 and not conf_per_pixel:
            loss, gscale = conf_rate * (b / P), conf_rate / P
        else:
            loss, gscale = cc / P, 1.0 / P
        loss = loss * ramp
        dls = g * (w * cw if conf_per_pixel else w)
        self.launches += 2
        out4 = torch.tensor([float(loss), float(conf_rate), float(gscale * ramp * cons_weight), float(loss * cons_weight)])
        return out4, dls.contiguous()

    def consistency(self, l0, l1, ls, m, lmask, loss_fn, conf_thresh, conf_per_pixel, ramp, cons_weight, dls=None):
        n, c, h, w = ls.shape
        lt = _mix1(l0, l1, m) if l1 is not None else l0
        pt = F.softmax(lt, dim=1)
        conf = (pt.max(1, keepdim=True)[0] >= conf_thresh).float() if conf_thresh > 0.0 else torch.ones((n, 1, h, w))
        q, g = self._q_and_grad(loss_fn, pt, lt, ls, c)
        wmask = lmask if lmask is not None else torch.ones((n, 1, h, w))
        return self._finish(q, g, wmask, conf, conf, conf_thresh, conf_per_pixel, ramp, cons_weight)

    def ict_consistency(self, l0, l1, ls, factors, lmask, loss_fn, conf_thresh, conf_per_pixel, ramp, cons_weight, dls=None):
        n, c, h, w = ls.shape
        f = factors.reshape(-1, 1, 1, 1)
        p0, p1 = F.softmax(l0, dim=1), F.softmax(l1, dim=1)
        pt, lt = _mix1(p0, p1, f), _mix1(l0, l1, f)
        if conf_thresh > 0.0:
            conf = (_mix1(p0.max(1, keepdim=True)[0], p1.max(1, keepdim=True)[0], f) >= conf_thresh).float()
        else:
            conf = torch.ones((n, 1, h, w))
        cw = conf
        if conf_per_pixel and conf_thresh > 0.0:          # b2_ict_conf_mean: batch-mean mask of the pixel
            cw = conf.mean(dim=0, keepdim=True).expand_as(conf)
            self.launches += 1
        q, g = self._q_and_grad(loss_fn, pt, lt, ls, c)
        wmask = lmask if lmask is not None else torch.ones((n, 1, h, w))
        return self._finish(q, g, wmask, conf, cw, conf_thresh, conf_per_pixel, ramp, cons_weight)


    # ------------------------------------------------------------------ augmentation consistency (aug_mt)
    @staticmethod
    def _affine_taps(theta, oh, ow, ih, iw):
        """affine_bilinear_tap() of losses.cu in tensor form: flat teacher offsets (N,4,OH,OW; -1 = outside) and weights
        in the order nw, ne, sw, se."""
        def lin(n):
            if n <= 1:
                return torch.zeros((max(n, 1),))
            i = torch.arange(n, dtype=torch.float32)
            step = torch.tensor(2.0 / (n - 1), dtype=torch.float32)
            return torch.where(i < n // 2, -1.0 + step * i, 1.0 - step * (n - 1 - i))
        xb = lin(ow)[None, None, :]; yb = lin(oh)[None, :, None]
        t = theta.reshape(-1, 6)
        gx = xb * t[:, 0, None, None] + yb * t[:, 1, None, None] + t[:, 2, None, None]
        gy = xb * t[:, 3, None, None] + yb * t[:, 4, None, None] + t[:, 5, None, None]
        ix = ((gx + 1.0) * 0.5) * float(iw - 1); iy = ((gy + 1.0) * 0.5) * float(ih - 1)
        fx, fy = torch.floor(ix), torch.floor(iy)
        we, ws = ix - fx, iy - fy
        ww, wn = 1.0 - we, 1.0 - ws
        x0 = fx.clamp(-2.0, float(iw)).long(); y0 = fy.clamp(-2.0, float(ih)).long()
        offs, wts = [], []
        for k, wk in enumerate((wn * ww, wn * we, ws * ww, ws * we)):
            xx, yy = x0 + (k & 1), y0 + (k >> 1)
            inside = (xx >= 0) & (xx < iw) & (yy >= 0) & (yy < ih)
            offs.append(torch.where(inside, yy * iw + xx, torch.full_like(xx, -1)))
            wts.append(wk)
        return torch.stack(offs, 1), torch.stack(wts, 1)

    @staticmethod
    def _gather_taps(x, offs, wts):
        """sum_k x[n, :, off_k] * w_k with zero for outside taps; x: (N,C,IH,IW) -> (N,C,OH,OW)."""
        n, c = x.shape[:2]
        flat = x.reshape(n, c, -1)
        out = torch.zeros((n, c) + tuple(offs.shape[2:]))
        for k in range(4):
            o = offs[:, k].reshape(n, 1, -1)
            v = torch.gather(flat, 2, o.clamp_min(0).expand(n, c, -1))
            v = torch.where(o >= 0, v, torch.zeros_like(v)).reshape(out.shape)
            out = out + v * wts[:, k][:, None]
        return out

    def affine_grid_sample(self, x, theta, out_hw=None):
        n, c, ih, iw = x.shape
        oh, ow = (ih, iw) if out_hw is None else out_hw
        offs, wts = self._affine_taps(theta, oh, ow, ih, iw)
        self.launches += 1
        return self._gather_taps(x, offs, wts)

    def aug_consistency(self, ltea, ls, theta, um0, um1, loss_fn, conf_thresh, conf_per_pixel, ramp, cons_weight, dls=None):
        n, c, h, w = ls.shape
        offs, wts = self._affine_taps(theta, h, w, h, w)
        lt = self._gather_taps(ltea, offs, wts)
        pt = self._gather_taps(F.softmax(ltea, dim=1), offs, wts)
        wmask = self._gather_taps(um0, offs, wts) * um1
        conf = (pt.max(1, keepdim=True)[0] >= conf_thresh).float() if conf_thresh > 0.0 else torch.ones((n, 1, h, w))
        q, g = self._q_and_grad(loss_fn, pt, lt, ls, c)
        return self._finish(q, g, wmask, conf, conf, conf_thresh, conf_per_pixel, ramp, cons_weight)


    # ------------------------------------------------------------------ VAT (csrc/vat.cu)
    def sample_l2norm(self, x):
        self.launches += 2
        return x.reshape(x.shape[0], -1).double().pow(2).sum(1).float().sqrt()

    def vat_adaptive_radius(self, x, vat_radius):
        self.launches += 2
        n = x.shape[0]
        dv = (x[:, :, 2:, :] - x[:, :, :-2, :]).reshape(n, -1).double().pow(2).sum(1).float()
        dh = (x[:, :, :, 2:] - x[:, :, :, :-2]).reshape(n, -1).double().pow(2).sum(1).float()
        return (torch.tensor(vat_radius, dtype=torch.float32) * torch.sqrt(dv + dh)) * 0.5

    def add_scaled_per_sample(self, x, e, mag, radius, out=None):
        self.launches += 1
        r = radius.view(-1, 1, 1, 1) if torch.is_tensor(radius) else torch.tensor(radius, dtype=torch.float32)
        d = (e / (mag.view(-1, 1, 1, 1) + 1e-12)) * r
        return d if x is None else x + d


    # ------------------------------------------------------------------ data-format boundary (csrc/input.cu)
    def normalize_to_tensor(self, img_u8, mean=None, std=None, out=None):
        self.launches += 1
        img = img_u8.numpy().astype(np.float64) * (1.0 / 255.0)
        if img.shape[3] == 4:
            alpha, img = img[..., 3:4], img[..., :3]
            if mean is not None:
                img = (img - np.asarray(mean, dtype=np.float64)[None, None, None, :] * alpha) / np.asarray(std, dtype=np.float64)
        elif mean is not None:
            img = (img - np.asarray(mean, dtype=np.float64)) / np.asarray(std, dtype=np.float64)
        return torch.from_numpy(np.ascontiguousarray(img.transpose(0, 3, 1, 2)).astype(np.float32))

    def labels_to_tensor(self, labels_u8):
        self.launches += 1
        return labels_u8[:, None].to(torch.int64)

    def mask_to_tensor(self, mask_u8):
        self.launches += 1
        return torch.from_numpy((mask_u8.numpy().astype(np.float64) * (1.0 / 255.0))[:, None].astype(np.float32))


class EmuEMA(object):
    """Stand-in for optim_weight_ema.EMAWeightOptimizer.step() on CPU modules (optim_weight_ema.py:21-25 arithmetic); the
    product class refuses CPU tensors by design."""

    def __init__(self, target_net, source_net, alpha):
        self.pairs = [(t, source_net.state_dict()[k]) for k, t in target_net.state_dict().items() if t.dtype == torch.float32]
        self.alpha = alpha
        with torch.no_grad():
            for t, s in self.pairs:
                t.copy_(s)

    def step(self):
        one_minus = 1.0 - self.alpha
        with torch.no_grad():
            for t, s in self.pairs:
                t.mul_(self.alpha)
                t.add_(s * one_minus)

"""Device side of the reference's DataLoader boundary.

`DeviceNormalizeToTensor(mean, std)` is the reference's `SegCVTransformNormalizeToTensor(mean, std)`
(datapipe/seg_transforms_cv.py:587-672) applied to a whole collated batch AFTER the host-to-device copy: the DataLoader
workers stop at uint8 arrays (`image_arr` HWC, `labels_arr`, `mask_arr`), the batch crosses PCIe as uint8 (4x fewer bytes for
images, 8x for labels) and the kernels of csrc/input.cu produce exactly the tensors the reference's collate function would
have produced on the host: `image` fp32 (N,3,H,W), `labels` int64 (N,1,H,W), `mask` fp32 (N,1,H,W) -- bit for bit (float64
arithmetic, one rounding to float32).
"""
import torch

from . import ops as O


class DeviceNormalizeToTensor(object):
    def __init__(self, mean, std):
        """mean / std: per-channel sequences (`seg_transforms.get_mean_std(ds, net)`: net.MEAN / net.STD) or None, None."""
        if (mean is None) != (std is None):
            raise ValueError('mean and std must be given together')
        self.mean = None if mean is None else [float(v) for v in mean]
        self.std = None if std is None else [float(v) for v in std]
        self.be = O.default_backend()

    def __call__(self, batch):
        """batch: dict with `image_arr` uint8 (N,H,W,3|4) and optionally `labels_arr` uint8 (N,H,W), `mask_arr` uint8 (N,H,W)
        (CUDA tensors, or pinned host tensors that are copied first).  Returns a new dict with `image` / `labels` / `mask`
        like the reference's transform (the *_arr entries are dropped, other entries pass through)."""
        dev = torch.device('cuda', torch.cuda.current_device())
        out = {k: v for k, v in batch.items() if k not in ('image_arr', 'labels_arr', 'mask_arr')}
        img = batch['image_arr']
        out['image'] = self.be.normalize_to_tensor(img if img.is_cuda else img.to(dev, non_blocking=True), self.mean, self.std)
        if 'labels_arr' in batch:
            lab = batch['labels_arr']
            out['labels'] = self.be.labels_to_tensor(lab if lab.is_cuda else lab.to(dev, non_blocking=True))
        if 'mask_arr' in batch:
            m = batch['mask_arr']
            out['mask'] = self.be.mask_to_tensor(m if m.is_cuda else m.to(dev, non_blocking=True))
        return out


class DeviceCropFlipNormalize(object):
    """Random crop (+ padding of small images) + random flips + normalise-to-tensor on the device: the chain
    `SegCVTransformRandomCrop(crop_size, crop_offset)` -> `SegCVTransformRandomFlip(hflip, vflip, hvflip)` ->
    `SegCVTransformNormalizeToTensor(mean, std)` of the reference's training pipelines (datapipe/seg_transforms_cv.py:102-167,
    445-520, 587-672; assembled in train_seg_semisup_mask_mt.py:150-179), as ONE gather kernel over the decoded uint8 images at
    their original sizes.

    The random parameters are drawn on the HOST with numpy, in the reference's order -- crop position `rng.uniform(0, 1, (2,))`
    (pairs: then `rng.uniform(-1, 1, (2,))` for the second crop's offset, clipped), flips `rng.binomial(1, 0.5, (3,))` (pairs:
    `(2, 3)`) -- from `crop_rng` / `flip_rng` (the reference's transforms own one RandomState each), so seeded runs select the
    same crops.  `draw_single` / `draw_pair` return plain dicts; `__call__` turns samples + parameters into the tensors the
    reference's collate function would have produced."""

    def __init__(self, crop_size, crop_offset=None, hflip=False, vflip=False, hvflip=False, mean=None, std=None, crop_rng=None,
                 flip_rng=None):
        import numpy as np
        if (mean is None) != (std is None):
            raise ValueError('mean and std must be given together')
        self.crop_size = np.array(crop_size)
        self.crop_offset = np.array([0, 0] if crop_offset is None else crop_offset)
        self.hflip, self.vflip, self.hvflip = bool(hflip), bool(vflip), bool(hvflip)
        if self.hvflip and self.crop_size[0] != self.crop_size[1]:
            raise ValueError('hvflip (transposition) needs a square crop')
        self.mean = None if mean is None else [float(v) for v in mean]
        self.std = None if std is None else [float(v) for v in std]
        self.crop_rng = crop_rng if crop_rng is not None else np.random.RandomState()
        self.flip_rng = flip_rng if flip_rng is not None else np.random.RandomState()
        self.be = None

    # ---- host: parameters, in the reference's draw order
    def _pad(self, img_hw):
        """(pad_top, pad_left, padded, padded size) of SegCVTransformPad.pad_single / pad_pair (:30-62)."""
        import numpy as np
        h, w = int(img_hw[0]), int(img_hw[1])
        if h < self.crop_size[0] or w < self.crop_size[1]:
            pad_h, pad_w = max(int(self.crop_size[0]) - h, 0), max(int(self.crop_size[1]) - w, 0)
            return pad_h // 2, pad_w // 2, 1, np.array([h + pad_h, w + pad_w])
        return 0, 0, 0, np.array([h, w])

    def draw_single(self, img_hw):
        import numpy as np
        top, left, padded, size = self._pad(img_hw)
        extra = size - self.crop_size
        pos = np.round(extra * self.crop_rng.uniform(0.0, 1.0, size=(2,))).astype(int)                       # :124-125
        on = np.array([self.hflip, self.vflip, self.hvflip])
        # the entry points add the flip transform only if a flip is enabled (train_seg_semisup_mask_mt.py:163-165)
        flags = ((self.flip_rng.binomial(1, 0.5, size=(3,)) != 0) & on) if on.any() else np.zeros(3, bool)              # :480-481
        return dict(pad_top=top, pad_left=left, padded=padded, pos=(int(pos[0]), int(pos[1])), flips=tuple(bool(f) for f in flags),
                    xf_cv=self._xf(pos, top, left, padded))

    def draw_pair(self, img_hw):
        import numpy as np
        top, left, padded, size = self._pad(img_hw)
        extra = size - self.crop_size
        pos0 = np.round(extra * self.crop_rng.uniform(0.0, 1.0, size=(2,))).astype(int)                      # :143-144
        pos1 = pos0 + np.round(self.crop_offset * self.crop_rng.uniform(-1.0, 1.0, size=(2,))).astype(int)   # :145
        pos1 = np.clip(pos1, np.array([0, 0]), extra)                                                        # :147
        on = np.array([[self.hflip, self.vflip, self.hvflip]])
        flags = ((self.flip_rng.binomial(1, 0.5, size=(2, 3)) != 0) & on) if on.any() else np.zeros((2, 3), bool)       # :503-504
        return tuple(dict(pad_top=top, pad_left=left, padded=padded, pos=(int(p[0]), int(p[1])), flips=tuple(bool(f) for f in fl),
                          xf_cv=self._xf(p, top, left, padded))
                     for p, fl in ((pos0, flags[0]), (pos1, flags[1])))

    @staticmethod
    def _xf(pos, pad_top, pad_left, padded):
        """The `xf_cv` entry the reference would leave in the sample for an identity input transform: translation by -pos
        (:128-132, :150-159) after the padding's translation (:56-60, :94-97)."""
        import numpy as np
        parts = [_mat_translation(-np.array(pos)[None, ::-1])]
        if padded:
            parts.append(_mat_translation(np.array([[pad_left, pad_top]])))
        return _mat_cat(*parts)[0]

    # ---- device
    @staticmethod
    def table(samples, params, crop_size):
        """numpy structured array of b2_crop_entry records (include/b200seg.h) for device-resident samples."""
        import numpy as np
        fields = [('image', 'u8'), ('labels', 'u8'), ('mask', 'u8'), ('h0', 'i4'), ('w0', 'i4'), ('pad_top', 'i4'), ('pad_left', 'i4'),
                  ('padded', 'i4'), ('pos_y', 'i4'), ('pos_x', 'i4'), ('crop_h', 'i4'), ('crop_w', 'i4'), ('flip_x', 'i4'),
                  ('flip_y', 'i4'), ('flip_d', 'i4')]
        arr = np.zeros(len(samples), dtype=np.dtype(fields, align=True))
        assert arr.dtype.itemsize == 72
        for i, (s, p) in enumerate(zip(samples, params)):
            img = s['image_arr']
            lab, msk = s.get('labels_arr'), s.get('mask_arr')
            arr[i] = (img.data_ptr(), 0 if lab is None else lab.data_ptr(), 0 if msk is None else msk.data_ptr(), img.shape[0],
                      img.shape[1], p['pad_top'], p['pad_left'], p['padded'], p['pos'][0], p['pos'][1], int(crop_size[0]),
                      int(crop_size[1]), int(p['flips'][0]), int(p['flips'][1]), int(p['flips'][2]))
        return arr

    def crops_u8(self, samples, params):
        """The cropped / flipped samples as pixels: (RGBA uint8 (N,h,w,4), labels int64 (N,1,h,w) | None, mask fp32 (N,1,h,w) | None)."""
        return self(samples, params, raw=True)

    def __call__(self, samples, params, colour=None, colour_params=None, raw=False):
        """samples: list of dicts with `image_arr` uint8 (H_i, W_i, 3) and optionally `labels_arr` / `mask_arr` uint8 (H_i, W_i),
        contiguous CUDA tensors (or host tensors, copied first); params: one `draw_*` dict per sample.  Returns a dict with
        `image` fp32 (N,3,h,w) and, if every sample has them, `labels` int64 (N,1,h,w) / `mask` fp32 (N,1,h,w).
        `colour` (a DeviceColourJitter) + `colour_params` (one `draw()` dict per sample): the strong-colour branch -- the crops
        stay uint8 RGBA, are jittered in place and standardised afterwards (crop -> flip -> colour -> normalise, the reference's
        order)."""
        if self.be is None:
            self.be = O.default_backend()
        dev = torch.device('cuda', torch.cuda.current_device())
        moved = []
        for s in samples:
            d = {}
            for k in ('image_arr', 'labels_arr', 'mask_arr'):
                if s.get(k) is not None:
                    t = s[k]
                    if t.dtype != torch.uint8:
                        raise ValueError('{} must be uint8'.format(k))
                    d[k] = (t if t.is_cuda else t.to(dev, non_blocking=True)).contiguous()
            if d['image_arr'].dim() != 3 or d['image_arr'].shape[2] != 3:
                raise ValueError('image should have 3 channels, not {}'.format(tuple(d['image_arr'].shape)))       # :654
            moved.append(d)
        tab = torch.from_numpy(self.table(moved, params, self.crop_size).view('u1').copy()).to(dev, non_blocking=True)
        want_labels = all('labels_arr' in d for d in moved)
        want_mask = all('mask_arr' in d for d in moved)
        h, w = int(self.crop_size[0]), int(self.crop_size[1])
        if colour is not None or raw:
            rgba, labels, mask = self.be.crop_flip_u8(tab, len(moved), h, w, want_labels, want_mask, dev)
            self._keep = (moved, tab)
            if raw:
                return rgba, labels, mask
            colour(rgba, colour_params)
            image = self.be.normalize_to_tensor(rgba, self.mean, self.std)
        else:
            image, labels, mask = self.be.crop_flip_normalize(tab, len(moved), h, w, self.mean, self.std, want_labels, want_mask, dev)
        self._keep = (moved, tab)          # inputs of the asynchronous launch stay alive until the next call
        out = {'image': image}
        if labels is not None:
            out['labels'] = labels
        if mask is not None:
            out['mask'] = mask
        return out


class DeviceColourJitter(object):
    """The reference's strong colour augmentation on the device: `tvt.Compose([tvt.RandomApply([tvt.ColorJitter(
    brightness, contrast, saturation, hue)], p), tvt.RandomGrayscale(grey_p)])` applied to the second sample of every unsupervised
    pair by SegCVTransformTVT (train_seg_semisup_mask_mt.py:169-179, datapipe/seg_transforms_cv.py:541-586).

    `draw()` consumes torch's global generator exactly like torchvision 0.26 does for one image -- RandomApply `torch.rand(1)`;
    ColorJitter.get_params `torch.randperm(4)` then one `uniform_` per enabled factor in the order brightness, contrast,
    saturation, hue; RandomGrayscale `torch.rand(1)` -- so a seeded run jitters the same way; `__call__` applies the drawn
    parameters to uint8 crops (N,H,W,3|4) in place with csrc/input.cu's kernels, byte-identical to Pillow."""
    BRIGHTNESS, CONTRAST, SATURATION, HUE = 0, 1, 2, 3

    def __init__(self, brightness=0.4, contrast=0.4, saturation=0.4, hue=0.1, p=0.8, grey_p=0.2):
        def rng_of(value, center=1.0, bound=(0.0, float('inf')), clip_first=True):
            # torchvision ColorJitter._check_input for a scalar
            if value < 0:
                raise ValueError('colour-jitter magnitudes must be non-negative')
            lo, hi = center - float(value), center + float(value)
            if clip_first:
                lo = max(lo, 0.0)
            if not bound[0] <= lo <= hi <= bound[1]:
                raise ValueError('colour-jitter range out of bounds')
            return None if lo == hi == center else (lo, hi)
        self.ranges = [rng_of(brightness), rng_of(contrast), rng_of(saturation),
                       rng_of(hue, center=0.0, bound=(-0.5, 0.5), clip_first=False)]
        self.p, self.grey_p = float(p), float(grey_p)
        self.be = None

    def draw(self):
        """Parameters for ONE image: dict(ops=[(op, factor), ...] in application order, grey=bool)."""
        ops = []
        if not (self.p < float(torch.rand(1))):                      # RandomApply.forward: `if self.p < torch.rand(1): return img`
            order = torch.randperm(4)                                 # ColorJitter.get_params
            fac = [None if r is None else float(torch.empty(1).uniform_(r[0], r[1])) for r in self.ranges]
            for fn_id in order.tolist():                              # ColorJitter.forward
                if fac[fn_id] is not None:
                    ops.append((fn_id, fac[fn_id]))
        grey = bool(float(torch.rand(1)) < self.grey_p)               # RandomGrayscale.forward
        return dict(ops=ops, grey=grey)

    @staticmethod
    def table(params):
        """numpy structured array of b2_colour_entry records (include/b200seg.h)."""
        import numpy as np
        dt = np.dtype([('n_ops', 'i4'), ('op', 'i4', (4,)), ('factor', 'f4', (4,)), ('hue_shift', 'i4', (4,)), ('grey', 'i4')])
        assert dt.itemsize == 56
        arr = np.zeros(len(params), dtype=dt)
        for i, p in enumerate(params):
            arr[i]['n_ops'] = len(p['ops'])
            for k, (op, fac) in enumerate(p['ops']):
                arr[i]['op'][k] = op
                if op == DeviceColourJitter.HUE:
                    arr[i]['hue_shift'][k] = int(np.int32(fac * 255).astype(np.uint8))      # _functional_pil.adjust_hue
                else:
                    arr[i]['factor'][k] = np.float32(fac)                                    # Image.blend takes a C float
            arr[i]['grey'] = int(p['grey'])
        return arr

    def __call__(self, images_u8, params):
        if self.be is None:
            self.be = O.default_backend()
        return self.be.colour_jitter(images_u8, self.table(params))


# ------------------------------------------------------------------------------------------------------------------------
# Scale / rotation crops: SegCVTransformRandomCropScaleHung (datapipe/seg_transforms_cv.py:169-303, the
# Pascal recipes' --aug_scale_hung) and SegCVTransformRandomCropRotateScale (:305-449, the ISIC recipes' --aug_max_scale /
# --aug_rot_mag).  The reference calls cv2.resize / cv2.warpAffine on uint8 arrays; OpenCV's uint8 paths are fixed-point integer
# algorithms (pinned 3.4.2, unchanged in the installed 4.13: tests/test_geom_pipeline.py compares the restatement with cv2 itself),
# so the host only prepares small integer tables per sample -- coefficient tables of the resize, the inverted matrix in 10-bit
# fixed point of the warp -- and csrc/input.cu (geom_u8_kernel) does the gather + interpolation in integer arithmetic.

def _mat_translation(xlats_xy):
    """datapipe/affine.py:73-84: (N,2,3) float32 translation matrices."""
    import numpy as np
    xf = np.zeros((len(xlats_xy), 2, 3), dtype=np.float32)
    xf[:, 0, 0] = xf[:, 1, 1] = 1.0
    xf[:, :, 2] = xlats_xy
    return xf


def _mat_scale(scale_xy):
    """datapipe/affine.py:86-97."""
    import numpy as np
    xf = np.zeros((len(scale_xy), 2, 3), dtype=np.float32)
    xf[:, 0, 0] = scale_xy[:, 0]
    xf[:, 1, 1] = scale_xy[:, 1]
    return xf


def _mat_rotation(thetas):
    """datapipe/affine.py:99-121: counter-clockwise, +y down."""
    import numpy as np
    c, s = np.cos(thetas), np.sin(thetas)
    xf = np.zeros((len(thetas), 2, 3), dtype=np.float32)
    xf[:, 0, 0] = xf[:, 1, 1] = c
    xf[:, 1, 0] = -s
    xf[:, 0, 1] = s
    return xf


def _mat_cat(*x):
    """datapipe/affine.py:43-71: x[0] . x[1] . ... (same numpy calls on the same dtypes, so the float32 products round alike)."""
    import numpy as np
    y = x[0]
    for b in x[1:]:
        a2, b2 = y[:, :, :2], b[:, :, :2]
        y = np.append(np.matmul(a2, b2), y[:, :, 2:3] + np.matmul(a2, b[:, :, 2:3]), axis=2)
    return y


def _mat_identity(n):
    """datapipe/affine.py:3-11."""
    import numpy as np
    xf = np.zeros((n, 2, 3), dtype=np.float32)
    xf[:, 0, 0] = xf[:, 1, 1] = 1.0
    return xf


def _mat_inv(m):
    """datapipe/affine.py:14-41 (inv_nx2x2 + inv_nx2x3), same numpy operations."""
    import numpy as np
    x = m[:, :, :2]
    rdet = 1.0 / (x[:, 0, 0] * x[:, 1, 1] - x[:, 1, 0] * x[:, 0, 1])
    y = np.zeros_like(x)
    y[:, 0, 0] = x[:, 1, 1] * rdet
    y[:, 1, 1] = x[:, 0, 0] * rdet
    y[:, 0, 1] = -x[:, 0, 1] * rdet
    y[:, 1, 0] = -x[:, 1, 0] * rdet
    return np.append(y, np.matmul(y, -m[:, :, 2:3]), axis=2)


def _mat_cv_to_torch(mtx, dst_size):
    """datapipe/affine.py:185-232 with src_size = None: OpenCV pixel-space matrices -> F.affine_grid matrices (align_corners)."""
    sx, sy = float(dst_size[1] - 1) / 2.0, float(dst_size[0] - 1) / 2.0
    n = len(mtx)
    mtx = _mat_inv(mtx)
    torch_cv = _mat_identity(n)
    torch_cv[:, 0, 0] = sx; torch_cv[:, 1, 1] = sy; torch_cv[:, 0, 2] = sx; torch_cv[:, 1, 2] = sy
    cv_torch = _mat_identity(n)
    cv_torch[:, 0, 0] = 1.0 / sx; cv_torch[:, 1, 1] = 1.0 / sy; cv_torch[:, 0, 2] = -1.0; cv_torch[:, 1, 2] = -1.0
    return _mat_cat(cv_torch, mtx, torch_cv)


def pair_flip_matrices(flags0, flags1, size_hw0, size_hw1, xf0, xf1):
    """SegCVTransformRandomFlip.transform_pair's update of `xf_cv` (seg_transforms_cv.py:506-525): sizes are those of the FLIPPED
    crops (the script reads them after flip_image)."""
    import numpy as np
    fl = np.array([flags0, flags1], dtype=bool)
    scale_xy = fl[:, :2] * -2 + 1
    xlat_xy = fl[:, :2] * (np.array([tuple(size_hw0)[::-1], tuple(size_hw1)[::-1]]).astype(float) - 1)
    hv = _mat_identity(2)
    hv[fl[:, 2]] = hv[fl[:, 2], ::-1, :]
    out = _mat_cat(hv, _mat_translation(xlat_xy), _mat_scale(scale_xy), np.stack([xf0, xf1], axis=0))
    return out[0], out[1]


NEAREST, LINEAR, AREA2 = 0, 1, 2


def resize_tables(src_hw, dst_hw):
    """Integer tables of cv2.resize for uint8 (OpenCV imgproc/resize.cpp): per destination column / row the nearest source index
    (INTER_NEAREST: min(floor(d * scale), n - 1), scale = 1 / (dst / src) in double), the left / upper source index of the linear
    filter and its two 11-bit coefficients packed as a0 | a1 << 16 ((float)((d + 0.5) * scale - 0.5), floor, float remainder,
    saturate_cast<short>(c * 2048); columns clamp the index and zero the remainder at the borders, rows keep the remainder and clip
    the row index at use).  Returns (xn, xl, xa, yn, yl, yb) int32 arrays and `area2` (INTER_LINEAR of an exact 2 x decimation is
    computed as a 2 x 2 box average)."""
    import numpy as np
    (sh, sw), (dh, dw) = src_hw, dst_hw
    scale_x, scale_y = 1.0 / (dw / sw), 1.0 / (dh / sh)
    eps = np.finfo(np.float64).eps
    area2 = bool(abs(scale_x - 2) < eps and abs(scale_y - 2) < eps)

    def axis(n_dst, n_src, scale, clamp):
        d = np.arange(n_dst, dtype=np.float64)
        near = np.minimum(np.floor(d * scale).astype(np.int64), n_src - 1)
        f = ((d + 0.5) * scale - 0.5).astype(np.float32)
        si = np.floor(f).astype(np.int64)
        f = f - si.astype(np.float32)
        if clamp:
            lo = si < 0
            f = np.where(lo, np.float32(0), f); si = np.where(lo, 0, si)
            hi = si >= n_src - 1
            f = np.where(hi, np.float32(0), f); si = np.where(hi, n_src - 1, si)
        a0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int64)
        a1 = np.rint(f * np.float32(2048)).astype(np.int64)
        return near.astype(np.int32), si.astype(np.int32), (a0 | (a1 << 16)).astype(np.int32)
    xn, xl, xa = axis(dw, sw, scale_x, True)
    yn, yl, yb = axis(dh, sh, scale_y, False)
    return (xn, xl, xa, yn, yl, yb), area2


def warp_tables(matrix_2x3, dst_hw):
    """Integer tables of cv2.warpAffine (OpenCV imgproc/imgwarp.cpp): the forward matrix is inverted in double, then
    adelta[x] = round(M00 * x * 1024), bdelta[x] = round(M10 * x * 1024), X0[y] = round((M01 * y + M02) * 1024), Y0[y] likewise
    (saturate_cast<int> = round half to even); the kernel adds the interpolation's rounding constant and shifts."""
    import numpy as np
    dh, dw = dst_hw
    m = np.asarray(matrix_2x3, dtype=np.float64).copy().ravel()
    det = m[0] * m[4] - m[1] * m[3]
    det = 1.0 / det if det != 0 else 0.0
    a11, a22 = m[4] * det, m[0] * det
    m[0] = a11; m[1] *= -det; m[3] *= -det; m[4] = a22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2] = b1; m[5] = b2
    xs, ys = np.arange(dw, dtype=np.float64), np.arange(dh, dtype=np.float64)

    def sat(v):
        return np.clip(np.rint(v), -2147483648, 2147483647).astype(np.int64).astype(np.int32)
    return sat(m[0] * xs * 1024), sat(m[3] * xs * 1024), sat((m[1] * ys + m[2]) * 1024), sat((m[4] * ys + m[5]) * 1024)


class _DeviceGeomBase(object):
    """Shared device side: table upload, geom_u8_kernel, optional colour jitter, normalise-to-tensor."""

    def _init_common(self, crop_size, crop_offset, hflip, vflip, hvflip, mean, std, rng, flip_rng):
        import numpy as np
        if (mean is None) != (std is None):
            raise ValueError('mean and std must be given together')
        self.crop_size = tuple(int(v) for v in crop_size)
        self.crop_size_arr = np.array(crop_size)
        self.crop_offset = np.array([0, 0] if crop_offset is None else crop_offset)
        self.hflip, self.vflip, self.hvflip = bool(hflip), bool(vflip), bool(hvflip)
        if self.hvflip and self.crop_size[0] != self.crop_size[1]:
            raise ValueError('hvflip (transposition) needs a square crop')
        self.mean = None if mean is None else [float(v) for v in mean]
        self.std = None if std is None else [float(v) for v in std]
        self.rng = rng if rng is not None else np.random.RandomState()
        self.flip_rng = flip_rng if flip_rng is not None else np.random.RandomState()
        self.be = None

    def _flips(self, pair):
        """SegCVTransformRandomFlip.transform_single / transform_pair (:478-481, :500-504)."""
        import numpy as np
        on = np.array([self.hflip, self.vflip, self.hvflip])
        if not on.any():              # the entry points add the flip transform only if a flip is enabled (:163-165): nothing is drawn
            return ((False,) * 3, (False,) * 3) if pair else (False,) * 3
        if not pair:
            return tuple(bool(f) for f in ((self.flip_rng.binomial(1, 0.5, size=(3,)) != 0) & on))
        fl = (self.flip_rng.binomial(1, 0.5, size=(2, 3)) != 0) & on[None]
        return tuple(bool(f) for f in fl[0]), tuple(bool(f) for f in fl[1])

    @staticmethod
    def entry_dtype():
        import numpy as np
        dt = np.dtype([('image', 'u8'), ('labels', 'u8'), ('mask', 'u8'), ('h0', 'i4'), ('w0', 'i4'), ('mode', 'i4'),
                       ('pad_top', 'i4'), ('pad_left', 'i4'), ('padded', 'i4'), ('pos_y', 'i4'), ('pos_x', 'i4'), ('src_h', 'i4'),
                       ('src_w', 'i4'), ('image_interp', 'i4'), ('mask_interp', 'i4'), ('tab_off', 'i4'), ('flip_x', 'i4'),
                       ('flip_y', 'i4'), ('flip_d', 'i4')], align=True)
        assert dt.itemsize == 88
        return dt

    def tables(self, samples, params):
        """(entries: structured array of b2_geom_entry records, tables: int32 array) for device-resident samples."""
        import numpy as np
        h, w = self.crop_size
        per = 3 * (h + w)
        ent = np.zeros(len(samples), dtype=self.entry_dtype())
        tab = np.zeros(len(samples) * per, dtype=np.int32)
        for i, (s, p) in enumerate(zip(samples, params)):
            img, lab, msk = s['image_arr'], s.get('labels_arr'), s.get('mask_arr')
            e = ent[i]
            e['image'] = img.data_ptr(); e['labels'] = 0 if lab is None else lab.data_ptr(); e['mask'] = 0 if msk is None else msk.data_ptr()
            e['h0'], e['w0'] = int(img.shape[0]), int(img.shape[1])
            e['mode'] = p['mode']
            e['tab_off'] = i * per
            e['flip_x'], e['flip_y'], e['flip_d'] = (int(f) for f in p['flips'])
            t = tab[i * per:(i + 1) * per]
            if p['mode'] == 0:
                e['pad_top'], e['pad_left'], e['padded'] = p['pad_top'], p['pad_left'], p['padded']
                e['pos_y'], e['pos_x'] = p['pos']
                e['src_h'], e['src_w'] = p['src_size']
                (xn, xl, xa, yn, yl, yb), area2 = resize_tables(p['src_size'], (h, w))
                t[0:w] = xn; t[w:2 * w] = xl; t[2 * w:3 * w] = xa
                t[3 * w:3 * w + h] = yn; t[3 * w + h:3 * w + 2 * h] = yl; t[3 * w + 2 * h:3 * w + 3 * h] = yb
                e['image_interp'] = AREA2 if (p['image_interp'] == LINEAR and area2) else p['image_interp']
                e['mask_interp'] = AREA2 if (p['mask_interp'] == LINEAR and area2) else p['mask_interp']
            else:
                ad, bd, x0, y0 = warp_tables(p['matrix'], (h, w))
                t[0:w] = ad; t[w:2 * w] = bd; t[3 * w:3 * w + h] = x0; t[3 * w + h:3 * w + 2 * h] = y0
                e['image_interp'], e['mask_interp'] = p['image_interp'], p['mask_interp']
        return ent, tab

    def crops_u8(self, samples, params):
        """The scaled / rotated / flipped crops as pixels: (RGBA uint8 (N,h,w,4), labels int64 | None, mask fp32 | None)."""
        return self(samples, params, raw=True)

    def __call__(self, samples, params, colour=None, colour_params=None, raw=False):
        """samples: list of dicts with `image_arr` uint8 (H_i, W_i, 3) [+ `labels_arr` / `mask_arr` uint8 (H_i, W_i)] (CUDA tensors,
        or host tensors that are copied first); params: one `draw_*` dict per sample.  Returns `image` fp32 (N,3,h,w) [+ `labels`
        int64 (N,1,h,w), `mask` fp32 (N,1,h,w)] as the reference's collate function would have produced them after
        scale / rotate crop -> flip -> [colour jitter] -> normalise."""
        if self.be is None:
            self.be = O.default_backend()
        dev = torch.device('cuda', torch.cuda.current_device())
        moved = []
        for s in samples:
            d = {}
            for k in ('image_arr', 'labels_arr', 'mask_arr'):
                if s.get(k) is not None:
                    t = s[k]
                    if t.dtype != torch.uint8:
                        raise ValueError('{} must be uint8'.format(k))
                    d[k] = (t if t.is_cuda else t.to(dev, non_blocking=True)).contiguous()
            if d['image_arr'].dim() != 3 or d['image_arr'].shape[2] != 3:
                raise ValueError('image should have 3 channels, not {}'.format(tuple(d['image_arr'].shape)))
            moved.append(d)
        ent, tab = self.tables(moved, params)
        ent_dev = torch.from_numpy(ent.view('u1').copy()).to(dev, non_blocking=True)
        tab_dev = torch.from_numpy(tab).to(dev, non_blocking=True)
        want_labels = all('labels_arr' in d for d in moved)
        want_mask = all('mask_arr' in d for d in moved)
        h, w = self.crop_size
        rgba, labels, mask = self.be.geom_u8(ent_dev, tab_dev, len(moved), h, w, want_labels, want_mask, dev)
        self._keep = (moved, ent_dev, tab_dev)
        if raw:
            return rgba, labels, mask
        if colour is not None:
            colour(rgba, colour_params)
        image = self.be.normalize_to_tensor(rgba, self.mean, self.std)
        self._keep = (moved, ent_dev, tab_dev)
        out = {'image': image}
        if labels is not None:
            out['labels'] = labels
        if mask is not None:
            out['mask'] = mask
        return out


class DeviceRandomCropScaleHung(_DeviceGeomBase):
    """`SegCVTransformRandomCropScaleHung(crop_size, crop_offset, uniform_scale)` -> `SegCVTransformRandomFlip` ->
    [`SegCVTransformTVT`] -> `SegCVTransformNormalizeToTensor` on the device.  Parameters are drawn on the host in the reference's
    order (`rng.randint(0, 11, (scale_dim,))`, `rng.uniform(0, 1, (2,))` [pairs: `rng.uniform(-1, 1, (2,))`]; flips from the flip
    transform's own generator)."""

    def __init__(self, crop_size, crop_offset=None, uniform_scale=True, hflip=False, vflip=False, hvflip=False, mean=None, std=None,
                 rng=None, flip_rng=None):
        self._init_common(crop_size, crop_offset, hflip, vflip, hvflip, mean, std, rng, flip_rng)
        self.uniform_scale = bool(uniform_scale)

    @staticmethod
    def _pad_to(img_hw, min_size):
        """SegCVTransformPad.pad_single / pad_pair (:30-100): (pad_top, pad_left, padded, padded size)."""
        import numpy as np
        h, w = int(img_hw[0]), int(img_hw[1])
        if h < min_size[0] or w < min_size[1]:
            pad_h, pad_w = max(int(min_size[0]) - h, 0), max(int(min_size[1]) - w, 0)
            return pad_h // 2, pad_w // 2, 1, np.array([h + pad_h, w + pad_w])
        return 0, 0, 0, np.array([h, w])

    def _xf(self, pos, sc_size, pad_top, pad_left, padded):
        """The `xf_cv` entry the reference would leave in the sample for an identity input transform (:216-229, :272-298)."""
        import numpy as np
        scale_yx = self.crop_size_arr / sc_size
        xlat_yx = (scale_yx - 1.0) * 0.5
        parts = [_mat_translation(xlat_yx[None, ::-1].astype(float)), _mat_scale(scale_yx[None, ::-1]),
                 _mat_translation(-np.array(pos)[None, ::-1].astype(float))]
        if padded:
            parts.append(_mat_translation(np.array([[pad_left, pad_top]])))
        return _mat_cat(*parts)[0]

    def draw_single(self, img_hw):
        import numpy as np
        scale_dim = 1 if self.uniform_scale else 2
        f_scale = 0.5 + self.rng.randint(0, 11, size=(scale_dim,)) / 10.0                                    # :199
        sc_size = np.round(self.crop_size_arr / f_scale).astype(int)                                         # :202
        top, left, padded, size = self._pad_to(img_hw, sc_size)                                              # :204
        extra = size - sc_size
        pos = np.round(extra * self.rng.uniform(0.0, 1.0, size=(2,))).astype(int)                            # :208-209
        return dict(mode=0, pad_top=top, pad_left=left, padded=padded, pos=(int(pos[0]), int(pos[1])),
                    src_size=(int(sc_size[0]), int(sc_size[1])), image_interp=LINEAR, mask_interp=LINEAR,   # :213-221
                    flips=self._flips(False), xf_cv=self._xf(pos, sc_size, top, left, padded))

    def draw_pair(self, img_hw):
        import numpy as np
        scale_dim = 1 if self.uniform_scale else 2
        f_scale1 = 0.5 + self.rng.randint(0, 11, size=(scale_dim,)) / 10.0                                   # :240
        sc_size1 = np.round(self.crop_size_arr / f_scale1).astype(int)                                       # :243
        max_sc = np.maximum(self.crop_size_arr, sc_size1)                                                    # :245
        top, left, padded, size = self._pad_to(img_hw, max_sc)                                               # :248
        extra = size - max_sc
        pos0 = np.round(extra * self.rng.uniform(0.0, 1.0, size=(2,))).astype(int)                           # :252
        pos1 = pos0 + np.round(self.crop_offset * self.rng.uniform(-1.0, 1.0, size=(2,))).astype(int)        # :253
        pos1 = np.clip(pos1, np.array([0, 0]), extra)                                                        # :255
        centre0, centre1 = pos0 + max_sc * 0.5, pos1 + max_sc * 0.5
        pos0 = np.round(centre0 - self.crop_size_arr * 0.5).astype(int)                                      # :260
        pos1 = np.round(centre1 - sc_size1 * 0.5).astype(int)                                                # :261
        f0, f1 = self._flips(True)
        crop = (int(self.crop_size_arr[0]), int(self.crop_size_arr[1]))
        p0 = dict(mode=0, pad_top=top, pad_left=left, padded=padded, pos=(int(pos0[0]), int(pos0[1])), src_size=crop,
                  image_interp=NEAREST, mask_interp=NEAREST, flips=f0,                                       # plain crop (:264)
                  xf_cv=self._xf(pos0, self.crop_size_arr, top, left, padded))
        p1 = dict(mode=0, pad_top=top, pad_left=left, padded=padded, pos=(int(pos1[0]), int(pos1[1])),
                  src_size=(int(sc_size1[0]), int(sc_size1[1])), image_interp=LINEAR, mask_interp=NEAREST, flips=f1,   # :267, :273
                  xf_cv=self._xf(pos1, sc_size1, top, left, padded))
        return p0, p1


class DeviceRandomCropRotateScale(_DeviceGeomBase):
    """`SegCVTransformRandomCropRotateScale(crop_size, crop_offset, rot_mag, max_scale, uniform_scale, constrain_rot_scale)` ->
    `SegCVTransformRandomFlip` -> [`SegCVTransformTVT`] -> `SegCVTransformNormalizeToTensor` on the device: cv2.warpAffine with
    BORDER_REFLECT_101 for the image and BORDER_CONSTANT for labels (255) / mask (0); nearest-neighbour sampling whenever the
    sample carries labels (:361-364, :424)."""

    def __init__(self, crop_size, crop_offset=None, rot_mag=0.0, max_scale=1.0, uniform_scale=True, constrain_rot_scale=True,
                 hflip=False, vflip=False, hvflip=False, mean=None, std=None, rng=None, flip_rng=None):
        import math
        import numpy as np
        self._init_common(crop_size, crop_offset, hflip, vflip, hvflip, mean, std, rng, flip_rng)
        self.rot_mag_rad = math.radians(rot_mag)
        self.log_max_scale = np.log(max_scale)
        self.uniform_scale, self.constrain_rot_scale = bool(uniform_scale), bool(constrain_rot_scale)

    def draw_single(self, img_hw, has_labels):
        import numpy as np
        lms = self.log_max_scale
        if self.uniform_scale:
            scale_yx = np.repeat(np.exp(self.rng.uniform(-lms, lms, size=(1,))), 2, axis=0)                  # :335-337
        else:
            scale_yx = np.exp(self.rng.uniform(-lms, lms, size=(2,)))                                        # :339
        theta = self.rng.uniform(-self.rot_mag_rad, self.rot_mag_rad, size=(1,))                             # :340
        sc_size = self.crop_size_arr / scale_yx                                                              # :343
        img_size = np.array([int(img_hw[0]), int(img_hw[1])])
        extra = np.maximum(img_size - sc_size, 0.0)
        centre = extra * self.rng.uniform(0.0, 1.0, size=(2,)) + np.minimum(sc_size, img_size) * 0.5         # :346-348
        xf = _mat_cat(_mat_translation(self.crop_size_arr[None, ::-1] * 0.5), _mat_rotation(theta),
                      _mat_scale(scale_yx[None, ::-1]), _mat_translation(-centre[None, ::-1]))               # :351-356
        if has_labels:
            interp = NEAREST                                                                                 # :361-362
        else:
            interp = int(self.rng.choice([NEAREST, LINEAR]))          # cv2.INTER_NEAREST = 0, cv2.INTER_LINEAR = 1 (:364)
        return dict(mode=1, matrix=xf[0], image_interp=interp, mask_interp=interp, flips=self._flips(False), xf_cv=xf[0])

    def draw_pair(self, img_hw, has_labels):
        import numpy as np
        lms, rot = self.log_max_scale, self.rot_mag_rad
        if self.constrain_rot_scale:                                                                         # :385-395
            if self.uniform_scale:
                scales = np.repeat(np.exp(self.rng.uniform(-lms, lms, size=(1, 1))), 2, axis=1)
            else:
                scales = np.exp(self.rng.uniform(-lms, lms, size=(1, 2)))
            thetas = self.rng.uniform(-rot, rot, size=(1,))
            scales = np.repeat(scales, 2, axis=0)
            thetas = np.repeat(thetas, 2, axis=0)
        else:                                                                                                # :396-402
            if self.uniform_scale:
                scales = np.repeat(np.exp(self.rng.uniform(-lms, lms, size=(2, 1))), 2, axis=1)
            else:
                scales = np.exp(self.rng.uniform(-lms, lms, size=(2, 2)))
            thetas = self.rng.uniform(-rot, rot, size=(2,))
        img_size = np.array([int(img_hw[0]), int(img_hw[1])])
        sc_size = self.crop_size_arr / scales.min(axis=0)                                                    # :407
        crop_centre = np.minimum(sc_size, img_size) * 0.5
        extra = np.maximum(img_size - sc_size, 0.0)
        centre0 = extra * self.rng.uniform(0.0, 1.0, size=(2,)) + crop_centre                                # :412
        offset1 = np.round(self.crop_offset * self.rng.uniform(-1.0, 1.0, size=(2,)))                        # :413
        centre_xlat = np.stack([centre0, centre0], axis=0)
        offset1_xlat = np.stack([np.zeros((2,)), offset1], axis=0)
        xfs = _mat_cat(_mat_translation(self.crop_size_arr[None, ::-1] * 0.5), _mat_translation(offset1_xlat[:, ::-1]),
                       _mat_rotation(thetas), _mat_scale(scales[:, ::-1]), _mat_translation(-centre_xlat[:, ::-1]))   # :418-424
        interp = NEAREST if has_labels else LINEAR                                                           # :427
        f0, f1 = self._flips(True)
        return tuple(dict(mode=1, matrix=xfs[i], image_interp=interp, mask_interp=interp, flips=f, xf_cv=xfs[i])
                     for i, f in ((0, f0), (1, f1)))


class DeviceTrainPipeline(object):
    """The train-time transform lists of the entry points (train_seg_semisup_mask_mt.py:147-179) assembled from the device
    transforms above: geometric stage by option -- `aug_scale_hung` -> SegCVTransformRandomCropScaleHung, `aug_max_scale != 1` or
    `aug_rot_mag != 0` -> SegCVTransformRandomCropRotateScale(constrain_rot_scale=True), else SegCVTransformRandomCrop (crop offset
    (0, 0)) --, SegCVTransformRandomFlip if any flip is enabled, and for the unsupervised samples with `aug_strong_colour`
    SegTransformToPair + SegCVTransformTVT(ColorJitter / RandomGrayscale) on the second member (`unsup_paired`), then
    SegCVTransformNormalizeToTensor(NET_MEAN, NET_STD).  As in the reference the supervised and the unsupervised list SHARE the
    geometric and flip transform objects (`train_unsup_transforms = train_transforms.copy()`, :167), i.e. their generators.

    The DataLoader side shrinks to decoding: `sup_batch` / `unsup_batch` take lists of uint8 samples (`image_arr` (H,W,3) +
    `labels_arr` for supervised, + `mask_arr` for unsupervised samples, as `ds_src.dataset(labels=..., mask=...)` yields them,
    :181-189) and return the tensors `seg_data.SegCollate` would have stacked."""

    def __init__(self, crop_size, mean, std, aug_hflip=False, aug_vflip=False, aug_hvflip=False, aug_scale_hung=False,
                 aug_max_scale=1.0, aug_scale_non_uniform=False, aug_rot_mag=0.0, aug_strong_colour=False,
                 aug_colour_brightness=0.4, aug_colour_contrast=0.4, aug_colour_saturation=0.4, aug_colour_hue=0.1,
                 aug_colour_prob=0.8, aug_colour_greyscale_prob=0.2, rng=None, flip_rng=None, script='mask_mt', aug_offset_range=16,
                 aug_free_scale_rot=False):
        """script='mask_mt': the lists above.  script='aug_mt': those of train_seg_semisup_aug_mt.py:126-163 -- the geometric
        transforms get `crop_offset = (aug_offset_range, aug_offset_range)` and `constrain_rot_scale = not aug_free_scale_rot`, and the
        unsupervised list is `[SegTransformToPair()] + train_transforms` (+ colour jitter on the second member): every unsupervised
        sample becomes a PAIR of differently augmented crops whose relative transform `xf0_to_1` (seg_data.SegCollate._compute_xf_0_to_1)
        drives the augmentation-consistency loss (`unsup_pair_batch`)."""
        if crop_size is None:
            raise NotImplementedError('the device pipeline needs a crop_size (fixed-size batches)')
        if script not in ('mask_mt', 'aug_mt'):
            raise ValueError('unknown script {!r}'.format(script))
        self.script = script
        offs = (aug_offset_range, aug_offset_range) if script == 'aug_mt' else (0, 0)
        any_flip = aug_hflip or aug_vflip or aug_hvflip
        common = dict(hflip=aug_hflip, vflip=aug_vflip, hvflip=aug_hvflip, mean=mean, std=std, flip_rng=flip_rng)
        if aug_scale_hung:                                                                                   # :151-152 / aug :130-131
            self.geom = DeviceRandomCropScaleHung(crop_size, offs, uniform_scale=not aug_scale_non_uniform, rng=rng, **common)
            self.kind = 'hung'
        elif aug_max_scale != 1.0 or aug_rot_mag != 0.0:                                                     # :153-155 / aug :132-135
            self.geom = DeviceRandomCropRotateScale(crop_size, offs, rot_mag=aug_rot_mag, max_scale=aug_max_scale,
                                                    uniform_scale=not aug_scale_non_uniform,
                                                    constrain_rot_scale=(not aug_free_scale_rot) if script == 'aug_mt' else True,
                                                    rng=rng, **common)
            self.kind = 'rot'
        else:                                                                                                # :156-157 / aug :136-137
            self.geom = DeviceCropFlipNormalize(crop_size, offs, crop_rng=rng, **common)
            self.kind = 'crop'
        self.any_flip = bool(any_flip)
        self.colour = DeviceColourJitter(aug_colour_brightness, aug_colour_contrast, aug_colour_saturation, aug_colour_hue,
                                         aug_colour_prob, aug_colour_greyscale_prob) if aug_strong_colour else None
        self.unsup_paired = bool(aug_strong_colour) or script == 'aug_mt'                                    # :170-179
        self.mean, self.std = self.geom.mean, self.geom.std

    def _draw(self, sample):
        hw = sample['image_arr'].shape[:2]
        if self.kind == 'rot':
            return self.geom.draw_single(hw, sample.get('labels_arr') is not None)
        return self.geom.draw_single(hw)

    def _crops(self, samples):
        params = [self._draw(s) for s in samples]
        return self.geom.crops_u8(samples, params)

    def sup_batch(self, samples):
        """-> dict(image fp32 (N,3,h,w), labels int64 (N,1,h,w))"""
        rgba, labels, _ = self._crops([{k: v for k, v in s.items() if k != 'mask_arr'} for s in samples])
        be = self.geom.be
        return dict(image=be.normalize_to_tensor(rgba, self.mean, self.std), labels=labels)

    def draw_pairs(self, samples):
        """Host side of `unsup_pair_batch`: per sample the two parameter dicts of the geometric transform's `transform_pair` with the
        flips' matrices folded into `xf_cv`, plus (N,2,3) `xf0_to_1_cv` / float32 `xf0_to_1` (SegCollate._compute_xf_0_to_1)."""
        import numpy as np
        h, w = (int(v) for v in self.geom.crop_size)
        params, xf01_cv = [], []
        for s in samples:
            hw = s['image_arr'].shape[:2]
            p0, p1 = self.geom.draw_pair(hw, s.get('labels_arr') is not None) if self.kind == 'rot' else self.geom.draw_pair(hw)
            if self.any_flip:
                # sizes of the flipped crops: a transposition swaps them (the crop is square when hvflip is enabled)
                sz0 = (w, h) if p0['flips'][2] else (h, w)
                sz1 = (w, h) if p1['flips'][2] else (h, w)
                x0, x1 = pair_flip_matrices(p0['flips'], p1['flips'], sz0, sz1, p0['xf_cv'], p1['xf_cv'])
                p0, p1 = dict(p0, xf_cv=x0), dict(p1, xf_cv=x1)
            params += [p0, p1]
            xf01_cv.append(_mat_cat(p1['xf_cv'][None], _mat_inv(p0['xf_cv'][None]))[0])
        xf01_cv = np.stack(xf01_cv, axis=0)
        xf01 = _mat_cv_to_torch(xf01_cv, (h, w)).astype(np.float32)
        return params, xf01_cv, xf01

    def unsup_pair_batch(self, samples):
        """script='aug_mt': -> dict(sample0=dict(image, mask), sample1=dict(image, mask), xf0_to_1 fp32 (N,2,3) device tensor,
        xf0_to_1_cv numpy): the tensors SegCollate stacks for train_seg_semisup_aug_mt.py:291-300 (`batch_ux0`, `batch_um0`,
        `batch_ux1`, `batch_um1`, `batch_ufx0_to_1`).  Colour jitter (aug_strong_colour) acts on the second member only."""
        assert self.script == 'aug_mt'
        use = [{k: v for k, v in s.items() if k != 'labels_arr'} for s in samples]
        params, xf01_cv, xf01 = self.draw_pairs(use)
        both = [s for s in use for _ in (0, 1)]
        rgba, _, mask = self.geom.crops_u8(both, params)
        be = self.geom.be
        rgba0, rgba1 = rgba[0::2].contiguous(), rgba[1::2].contiguous()
        if self.colour is not None:
            self.colour(rgba1, [self.colour.draw() for _ in use])
        dev = rgba.device
        return dict(sample0=dict(image=be.normalize_to_tensor(rgba0, self.mean, self.std), mask=mask[0::2].contiguous()),
                    sample1=dict(image=be.normalize_to_tensor(rgba1, self.mean, self.std), mask=mask[1::2].contiguous()),
                    xf0_to_1=torch.from_numpy(xf01).to(dev), xf0_to_1_cv=xf01_cv)

    def unsup_batch(self, samples):
        """-> dict(image, mask), or with aug_strong_colour dict(sample0=dict(image, mask), sample1=dict(image, mask)) where sample1 is
        the colour-jittered copy of the same crop (SegTransformToPair, then SegCVTransformTVT on the second member only)."""
        if self.script == 'aug_mt':
            return self.unsup_pair_batch(samples)
        rgba, _, mask = self._crops([{k: v for k, v in s.items() if k != 'labels_arr'} for s in samples])
        be = self.geom.be
        image0 = be.normalize_to_tensor(rgba, self.mean, self.std)
        if not self.unsup_paired:
            return dict(image=image0, mask=mask)
        cparams = [self.colour.draw() for _ in samples]
        rgba1 = rgba.clone()
        self.colour(rgba1, cparams)
        return dict(sample0=dict(image=image0, mask=mask), sample1=dict(image=be.normalize_to_tensor(rgba1, self.mean, self.std), mask=mask))

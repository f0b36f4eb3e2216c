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
        flags = (self.flip_rng.binomial(1, 0.5, size=(3,)) != 0) & np.array([self.hflip, self.vflip, self.hvflip])   # :480-481
        return dict(pad_top=top, pad_left=left, padded=padded, pos=(int(pos[0]), int(pos[1])), flips=tuple(bool(f) for f in flags))

    def draw_pair(self, img_hw):
        import numpy as np
        top, left, padded, size = self._pad(img_hw)
        extra = size - self.crop_size
        pos0 = np.round(extra * self.crop_rng.uniform(0.0, 1.0, size=(2,))).astype(int)                      # :143-144
        pos1 = pos0 + np.round(self.crop_offset * self.crop_rng.uniform(-1.0, 1.0, size=(2,))).astype(int)   # :145
        pos1 = np.clip(pos1, np.array([0, 0]), extra)                                                        # :147
        flags = (self.flip_rng.binomial(1, 0.5, size=(2, 3)) != 0) & np.array([[self.hflip, self.vflip, self.hvflip]])   # :503-504
        return tuple(dict(pad_top=top, pad_left=left, padded=padded, pos=(int(p[0]), int(p[1])), flips=tuple(bool(f) for f in fl))
                     for p, fl in ((pos0, flags[0]), (pos1, flags[1])))

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

    def __call__(self, samples, params, colour=None, colour_params=None):
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
        if colour is not None:
            rgba, labels, mask = self.be.crop_flip_u8(tab, len(moved), h, w, want_labels, want_mask, dev)
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

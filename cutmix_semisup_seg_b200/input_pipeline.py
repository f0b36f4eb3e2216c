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

    def __call__(self, samples, params):
        """samples: list of dicts with `image_arr` uint8 (H_i, W_i, 3) and optionally `labels_arr` / `mask_arr` uint8 (H_i, W_i),
        contiguous CUDA tensors (or host tensors, copied first); params: one `draw_*` dict per sample.  Returns a dict with
        `image` fp32 (N,3,h,w) and, if every sample has them, `labels` int64 (N,1,h,w) / `mask` fp32 (N,1,h,w)."""
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
        image, labels, mask = self.be.crop_flip_normalize(tab, len(moved), h, w, self.mean, self.std, want_labels, want_mask, dev)
        self._keep = (moved, tab)          # inputs of the asynchronous launch stay alive until the next call
        out = {'image': image}
        if labels is not None:
            out['labels'] = labels
        if mask is not None:
            out['mask'] = mask
        return out

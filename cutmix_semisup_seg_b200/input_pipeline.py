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

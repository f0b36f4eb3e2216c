"""Drop-in for the reference's `mask_gen` module (mask_gen.py:9-142): CutMix / CutOut box masks.

Same public surface — `MaskGenerator`, `BoxMaskGenerator(prop_range, n_boxes=1,
random_aspect_ratio=True, prop_by_area=True, within_bounds=True, invert=False)` with
`generate_params`, `torch_masks_from_params`, `append_to_batch`, and `AddMaskParamsToBatch` —
and the same numpy RNG draw order (props, aspect, positions), so seeded masks are identical.

B200-native addition: `generate_boxes` returns the 4 integers per box instead of a dense
(N,1,H,W) float64 mask, and `torch_masks_from_params` rasterises such compact params on the GPU
(`b2_box_mask_rasterize`), which removes the 16.8 MB/iteration host->device mask copy of the
reference loop (train_seg_semisup_mask_mt.py:331).  Sampling stays pure numpy: it runs inside
DataLoader worker processes (collate hook) and must never touch CUDA.
"""
import numpy as np
import torch


class MaskGenerator(object):
    """Abstract mask generator (reference mask_gen.py:9-23)."""

    def generate_params(self, n_masks, mask_shape, rng=None):
        raise NotImplementedError('Abstract')

    def append_to_batch(self, *batch):
        images = batch[0]
        return batch + (self.generate_params(len(images), images.shape[2:4]),)

    def torch_masks_from_params(self, t_params, mask_shape, torch_device):
        raise NotImplementedError('Abstract')


def _resolve_slices(edges, extent):
    """Clip float rectangle edges [lo, hi) exactly like numpy basic slicing `a[int(lo):int(hi)]` does
    (truncation toward zero, negative indices wrap, empty when stop <= start)."""
    lo = np.trunc(edges[..., 0]).astype(np.int64)
    hi = np.trunc(edges[..., 1]).astype(np.int64)
    out = np.zeros(edges.shape[:-1] + (2,), dtype=np.int32)
    flat_lo, flat_hi, flat_out = lo.reshape(-1), hi.reshape(-1), out.reshape(-1, 2)
    for j in range(flat_lo.shape[0]):
        start, stop, _ = slice(int(flat_lo[j]), int(flat_hi[j])).indices(extent)
        flat_out[j, 0] = start
        flat_out[j, 1] = max(stop, start)
    return out


class BoxMaskGenerator(MaskGenerator):
    """Random box masks (reference mask_gen.py:46-120)."""

    def __init__(self, prop_range, n_boxes=1, random_aspect_ratio=True, prop_by_area=True, within_bounds=True,
                 invert=False):
        if isinstance(prop_range, float):
            prop_range = (prop_range, prop_range)
        self.prop_range = prop_range
        self.n_boxes = n_boxes
        self.random_aspect_ratio = random_aspect_ratio
        self.prop_by_area = prop_by_area
        self.within_bounds = within_bounds
        self.invert = invert

    # -- sampling -----------------------------------------------------------------------------
    def _sample_rectangles(self, n_masks, mask_shape, rng):
        """Float rectangles (N, B, 4) = y0, x0, y1, x1, drawn in the reference's RNG order
        (mask_gen.py:73-108)."""
        if rng is None:
            rng = np.random
        lo, hi = self.prop_range
        shape_nb = (n_masks, self.n_boxes)
        box_scale = np.sqrt(1.0 / self.n_boxes)
        if self.prop_by_area:
            area = rng.uniform(lo, hi, size=shape_nb)
            degenerate = area == 0.0
            if self.random_aspect_ratio:
                frac_y = np.exp(rng.uniform(low=0.0, high=1.0, size=shape_nb) * np.log(area))
                frac_x = area / frac_y
                frac_y = frac_y * box_scale
                frac_x = frac_x * box_scale
            else:
                # The reference aliases y_props and x_props here (mask_gen.py:84) and then scales
                # "both" in place (:86-87), so the shared array is scaled twice.
                frac_y = (np.sqrt(area) * box_scale) * box_scale
                frac_x = frac_y.copy()
            frac_y[degenerate] = 0
            frac_x[degenerate] = 0
        else:
            if self.random_aspect_ratio:
                frac_y = rng.uniform(lo, hi, size=shape_nb)
                frac_x = rng.uniform(lo, hi, size=shape_nb)
                frac_y = frac_y * box_scale
                frac_x = frac_x * box_scale
            else:
                # same aliasing quirk as above (mask_gen.py:96-99): the single array is scaled twice
                frac_y = (rng.uniform(lo, hi, size=shape_nb) * box_scale) * box_scale
                frac_x = frac_y.copy()
        extent = np.array(mask_shape)
        sizes = np.round(np.stack([frac_y, frac_x], axis=2) * extent[None, None, :])
        u = rng.uniform(low=0.0, high=1.0, size=sizes.shape)
        if self.within_bounds:
            top_left = np.round((extent - sizes) * u)
            return np.concatenate([top_left, top_left + sizes], axis=2)
        centre = np.round(extent * u)
        return np.concatenate([centre - sizes * 0.5, centre + sizes * 0.5], axis=2)

    def generate_boxes(self, n_masks, mask_shape, rng=None):
        """Compact mask parameters: int32 (N, n_boxes, 4) = [y0, y1, x0, x1) half-open pixel ranges,
        already resolved with numpy slice semantics.  Rasterised by `torch_masks_from_params`."""
        rect = self._sample_rectangles(n_masks, tuple(mask_shape), rng)
        ys = _resolve_slices(np.stack([rect[..., 0], rect[..., 2]], axis=-1), mask_shape[0])
        xs = _resolve_slices(np.stack([rect[..., 1], rect[..., 3]], axis=-1), mask_shape[1])
        return np.concatenate([ys, xs], axis=-1).astype(np.int32)

    @staticmethod
    def rasterize_boxes_numpy(boxes, mask_shape, invert):
        """CPU rasterisation of `generate_boxes` output -> (N,1,H,W) float64 (toggle semantics,
        mask_gen.py:110-116)."""
        n = boxes.shape[0]
        masks = np.zeros((n, 1) + tuple(mask_shape)) if invert else np.ones((n, 1) + tuple(mask_shape))
        for i in range(n):
            for y0, y1, x0, x1 in boxes[i]:
                region = masks[i, 0, y0:y1, x0:x1]
                masks[i, 0, y0:y1, x0:x1] = 1 - region
        return masks

    def generate_params(self, n_masks, mask_shape, rng=None):
        """Reference-compatible: dense masks (N,1,H,W) float64 generated on the CPU."""
        boxes = self.generate_boxes(n_masks, mask_shape, rng)
        return self.rasterize_boxes_numpy(boxes, mask_shape, self.invert)

    # -- device side --------------------------------------------------------------------------
    def torch_masks_from_params(self, t_params, mask_shape, torch_device):
        """Dense params (N,1,H,W) pass through (reference mask_gen.py:119-120; the collate hook already made them float32,
        :141 -- a float64 array straight from `generate_params` is cast here, because the mix / loss kernels read raw fp32
        pointers); compact int32 box params (N,B,4) are rasterised on the GPU."""
        if t_params.dim() == 4:
            if t_params.dtype != torch.float32 or t_params.device != torch.device(torch_device):
                t_params = t_params.to(device=torch_device, dtype=torch.float32)
            return t_params
        if t_params.dim() != 3 or t_params.shape[-1] != 4:
            raise ValueError('mask params must be (N,1,H,W) masks or (N,n_boxes,4) boxes')
        from cutmix_semisup_seg_b200 import ops
        boxes = t_params.to(device=torch_device, dtype=torch.int32).contiguous()
        return ops.default_backend().box_mask_rasterize(boxes, int(mask_shape[0]), int(mask_shape[1]),
                                                        0.0 if self.invert else 1.0)


class AddMaskParamsToBatch(object):
    """Collate-time hook (reference mask_gen.py:123-142): attaches 'mask_params' to every sample.
    `compact=True` attaches the 4-int boxes instead of a dense float32 mask."""

    def __init__(self, mask_gen, compact=False):
        self.mask_gen = mask_gen
        self.compact = compact

    def __call__(self, batch):
        first = batch[0]
        ref_sample = first['sample0'] if 'sample0' in first else first
        mask_size = ref_sample['image'].shape[1:3]
        if self.compact:
            params = self.mask_gen.generate_boxes(len(batch), mask_size)
            for sample, p in zip(batch, params):
                sample['mask_params'] = p
        else:
            params = self.mask_gen.generate_params(len(batch), mask_size)
            for sample, p in zip(batch, params):
                sample['mask_params'] = p.astype(np.float32)
        return batch

"""Synthetic stand-ins for the reference's DataLoader batches (tensor contract of SURVEY.md §8c/§8d):
image fp32 (N,3,H,W) ~ N(0,1) (standardised images), labels int64 (N,1,H,W) with a 255 = ignore band,
valid masks fp32 (N,1,H,W) in [0,1] with a zero border and a fractional ramp column, and CutMix box
parameters from `mask_gen.BoxMaskGenerator` with a seeded numpy RandomState.  Used by bench.py, the smoke
test and the `--dataset synthetic` mode of the entry point (there is no network access for real data)."""
import math
from collections import OrderedDict

import numpy as np
import torch


def make_sup_batch(n, h, w, num_classes, seed, device='cpu', pin=False):
    g = torch.Generator().manual_seed(seed)
    image = torch.randn((n, 3, h, w), generator=g)
    labels = torch.randint(0, num_classes, (n, 1, h, w), generator=g, dtype=torch.int64)
    labels[:, :, :min(8, h // 4)] = 255
    if pin:
        image, labels = image.pin_memory(), labels.pin_memory()
    return image.to(device), labels.to(device)


def make_valid_mask(n, h, w):
    um = torch.ones((n, 1, h, w))
    b = min(8, h // 8, w // 8)
    if b > 0:
        um[:, :, :b] = 0; um[:, :, -b:] = 0; um[:, :, :, :b] = 0; um[:, :, :, -b:] = 0
    um[:, :, :, w // 2] *= 0.5
    return um


def make_unsup_batch(n, h, w, seed, mask_generator, mask_mix=True, compact_masks=True, paired=False, device='cpu',
                     pin=False):
    """Dict for MeanTeacherStep.step (mix mode: two views + mask params; cut mode: one view)."""
    g = torch.Generator().manual_seed(seed)
    rng = np.random.RandomState(12345 + seed)
    out = {}

    def view():
        tea = torch.randn((n, 3, h, w), generator=g)
        stu = tea + 0.1 * torch.randn((n, 3, h, w), generator=g) if paired else tea
        return tea, stu
    if mask_mix:
        out['ux0_tea'], out['ux0_stu'] = view()
        out['ux1_tea'], out['ux1_stu'] = view()
        out['um0'], out['um1'] = make_valid_mask(n, h, w), make_valid_mask(n, h, w)
    else:
        out['ux_tea'], out['ux_stu'] = view()
        out['um'] = make_valid_mask(n, h, w)
    if compact_masks:
        out['mask_params'] = torch.from_numpy(mask_generator.generate_boxes(n, (h, w), rng=rng))
    else:
        out['mask_params'] = torch.from_numpy(mask_generator.generate_params(n, (h, w), rng=rng).astype(np.float32))
    res = {}
    cache = {}
    for k, v in out.items():
        if id(v) not in cache:
            t = v.pin_memory() if pin else v
            cache[id(v)] = t.to(device)
        res[k] = cache[id(v)]
    return res


def make_ict_batch(n, h, w, seed, ict_alpha, paired=False, device='cpu', pin=False):
    """Dict for MeanTeacherStep.step in ICT mode (train_seg_semisup_ict.py:270-307): two unlabelled views with their valid
    masks and one Beta(ict_alpha, ict_alpha) mix factor per sample -- float64 numpy draws cast to float32 like the
    reference's `torch.tensor(..., dtype=torch.float)`."""
    g = torch.Generator().manual_seed(seed)
    rng = np.random.RandomState(12345 + seed)
    out = {}

    def view():
        tea = torch.randn((n, 3, h, w), generator=g)
        stu = tea + 0.1 * torch.randn((n, 3, h, w), generator=g) if paired else tea
        return tea, stu
    out['ux0_tea'], out['ux0_stu'] = view()
    out['ux1_tea'], out['ux1_stu'] = view()
    out['um0'], out['um1'] = make_valid_mask(n, h, w), make_valid_mask(n, h, w)
    out['ict_mix_factors'] = torch.tensor(rng.beta(ict_alpha, ict_alpha, size=(n,)), dtype=torch.float)
    res, cache = {}, {}
    for k, v in out.items():
        if id(v) not in cache:
            t = v.pin_memory() if pin else v
            cache[id(v)] = t.to(device)
        res[k] = cache[id(v)]
    return res


def make_aug_batch(n, h, w, seed, paired=False, rot_mag=10.0, max_scale=1.2, offset_range=4.0, device='cpu', pin=False):
    """Dict for MeanTeacherStep.step in augmentation-consistency mode (train_seg_semisup_aug_mt.py:275-281): two views of the
    unlabelled images with their valid masks and `xf0_to_1` (N,2,3) fp32, the affine map from the student's (view 1)
    normalised coordinates to the teacher's (view 0) that the reference's DataLoader derives from the two crops' transforms
    (datapipe/seg_data.py:223-231).  Each view draws its own rotation (+-rot_mag degrees), log-uniform scale in
    [1/max_scale, max_scale] and crop offset (+-offset_range pixels) like SegCVTransformRandomCropRotateScale
    (datapipe/seg_transforms_cv.py:310-341, 412); the map is the difference of the two."""
    g = torch.Generator().manual_seed(seed)
    rng = np.random.RandomState(12345 + seed)
    out = {}
    out['ux0'] = torch.randn((n, 3, h, w), generator=g)
    out['ux1'] = out['ux0'] + 0.1 * torch.randn((n, 3, h, w), generator=g) if paired else torch.randn((n, 3, h, w), generator=g)
    out['um0'], out['um1'] = make_valid_mask(n, h, w), make_valid_mask(n, h, w)
    rot, lms = np.radians(rot_mag), np.log(max_scale)
    ang = rng.uniform(-rot, rot, size=(n,)) - rng.uniform(-rot, rot, size=(n,))
    sc = np.exp(rng.uniform(-lms, lms, size=(n,)) - rng.uniform(-lms, lms, size=(n,)))
    off = np.round(offset_range * rng.uniform(-1.0, 1.0, size=(n, 2))) - np.round(offset_range * rng.uniform(-1.0, 1.0, size=(n, 2)))
    ty, tx = off[:, 0] * 2.0 / max(h - 1, 1), off[:, 1] * 2.0 / max(w - 1, 1)      # pixels -> normalised [-1, 1] coordinates
    asp = float(h) / float(w)              # normalised coordinates: a rotation in pixel space is sheared by the aspect ratio
    xf = np.zeros((n, 2, 3), dtype=np.float64)
    xf[:, 0, 0] = sc * np.cos(ang); xf[:, 0, 1] = -sc * np.sin(ang) * asp; xf[:, 0, 2] = tx
    xf[:, 1, 0] = sc * np.sin(ang) / asp; xf[:, 1, 1] = sc * np.cos(ang); xf[:, 1, 2] = ty
    out['xf0_to_1'] = torch.tensor(xf.astype(np.float32))
    res, cache = {}, {}
    for k, v in out.items():
        if id(v) not in cache:
            t = v.pin_memory() if pin else v
            cache[id(v)] = t.to(device)
        res[k] = cache[id(v)]
    return res


def make_vat_batch(n, h, w, seed, paired=False, with_noise=False, device='cpu', pin=False):
    """Dict for MeanTeacherStep.step in VAT mode (train_seg_semisup_vat_mt.py:364-380): one unlabelled view (or the weak /
    strong pair of `--aug_strong_colour`) with its valid mask; the key 'vat' marks the batch.  `with_noise`: also carry the
    N(0,1) draw behind the initial perturbation (reference: torch.randn on the device, :222) so that two implementations
    can be driven with the same draw."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    out['ux_tea'] = torch.randn((n, 3, h, w), generator=g)
    out['ux_stu'] = out['ux_tea'] + 0.1 * torch.randn((n, 3, h, w), generator=g) if paired else out['ux_tea']
    out['um'] = make_valid_mask(n, h, w)
    out['vat'] = torch.ones((1,))
    if with_noise:
        out['noise'] = torch.randn((n, 3, h, w), generator=g)
    res, cache = {}, {}
    for k, v in out.items():
        if id(v) not in cache:
            t = v.pin_memory() if pin else v
            cache[id(v)] = t.to(device)
        res[k] = cache[id(v)]
    return res


def condition_classifier(net, gain):
    """Scale the last classification layer so that teacher soft-max confidences straddle the 0.97 threshold
    on random inputs (random-init networks give conf_rate == 0, i.e. a vacuous consistency loss)."""
    with torch.no_grad():
        for name, p in net.named_parameters():
            if (name.startswith('layer5.') or 'classifier.classifier.6' in name or name.startswith('deeplab.classifier.4.')
                    or name.startswith('final_clf.')) and p.dim() == 4:
                p.mul_(gain)


def synth_state_dict(template, seed=0, logit_gain=1.0, final_keys=()):
    """Deterministic, well-conditioned synthetic weights for a state_dict-shaped mapping (random-init or
    checkpoint weights cannot be downloaded here): conv weights ~ N(0, 2/fan_in); BatchNorm gamma ~ U(0.5, 1.5) (U(0.15, 0.35)
    for the last BatchNorm of a residual unit and for down-sample BatchNorms, so that activations stay O(1) through 33
    residual units), beta ~ N(0, 0.1), running_mean ~ N(0, 0.1), running_var ~ U(0.5, 1.5), biases ~ N(0, 0.1); the weights
    named in `final_keys` are multiplied by `logit_gain`.  One CPU generator per tensor, seeded by (seed, position): the same
    values on every machine.  (oracle/torch_oracle.py carries an identical statement for the CPU side;
    tests/test_fullsize_recipe.py checks that the two agree bit for bit.)"""
    out = OrderedDict()
    for i, (k, v) in enumerate(template.items()):
        g = torch.Generator().manual_seed(seed * 100003 + i)
        shape = tuple(v.shape)
        if v.dtype != torch.float32:
            out[k] = torch.zeros(shape, dtype=v.dtype)
        elif k.endswith('running_mean'):
            out[k] = torch.randn(shape, generator=g) * 0.1
        elif k.endswith('running_var'):
            out[k] = torch.rand(shape, generator=g) + 0.5
        elif len(shape) == 4:
            w = torch.randn(shape, generator=g) * math.sqrt(2.0 / (shape[1] * shape[2] * shape[3]))
            out[k] = w * logit_gain if k in final_keys else w
        elif k.endswith('.weight'):
            small = k.endswith('bn3.weight') or k.endswith('downsample.1.weight')
            out[k] = torch.rand(shape, generator=g) * 0.2 + 0.15 if small else torch.rand(shape, generator=g) + 0.5
        else:
            out[k] = torch.randn(shape, generator=g) * 0.1
    return out


class U8ImageSource(object):
    """`--dataset synthetic_u8`: a seeded stand-in for a DECODED image data set -- uint8 images of assorted sizes around the crop
    size (some smaller than the crop: they get padded), uint8 label maps with a 255 = ignore band, and the all-valid 255 mask
    the reference's data sets hand out for unlabelled samples -- kept resident on the device.  The samples have the format of
    `ds_src.dataset(...)[i]` before the transform list (train_seg_semisup_mask_mt.py:181-189: `image_arr`, `labels_arr`,
    `mask_arr`), so the train-time pipeline (input_pipeline.DeviceTrainPipeline) runs on them exactly as it would on real decoded
    images.  The semi-supervised split follows the reference's scheme: a seeded permutation whose first `n_sup` indices are the
    supervised subset, the unsupervised subset is every sample (`n_unsup == -1`) or the first `n_unsup` (:103-118)."""

    def __init__(self, n_images, crop_hw, num_classes, seed, device, n_sup=100, n_unsup=-1, split_seed=12345):
        rs = np.random.RandomState(seed)
        h, w = int(crop_hw[0]), int(crop_hw[1])
        self.samples = []
        for i in range(n_images):
            f_h, f_w = rs.uniform(0.8, 1.7, size=2)
            ih, iw = max(8, int(round(h * f_h))), max(8, int(round(w * f_w)))
            # smooth-ish content (blocks of 8 x 8) so that interpolation and colour jitter act on structure, not on white noise
            coarse = rs.randint(0, 256, size=((ih + 7) // 8, (iw + 7) // 8, 3)).astype(np.uint8)
            img = np.repeat(np.repeat(coarse, 8, axis=0), 8, axis=1)[:ih, :iw]
            img = np.clip(img.astype(np.int16) + rs.randint(-12, 13, size=img.shape), 0, 255).astype(np.uint8)
            lab = np.repeat(np.repeat(rs.randint(0, num_classes, size=((ih + 15) // 16, (iw + 15) // 16)), 16, axis=0), 16, axis=1)[:ih, :iw]
            lab = lab.astype(np.uint8)
            lab[:max(1, ih // 16)] = 255
            self.samples.append(dict(image_arr=torch.from_numpy(np.ascontiguousarray(img)).to(device),
                                     labels_arr=torch.from_numpy(np.ascontiguousarray(lab)).to(device),
                                     mask_arr=torch.full((ih, iw), 255, dtype=torch.uint8, device=device)))
        perm = np.random.RandomState(split_seed).permutation(n_images)
        n_sup = n_images if n_sup == -1 else min(int(n_sup), n_images)
        self.sup_ndx = perm[:n_sup]
        self.unsup_ndx = perm if n_unsup == -1 else perm[:min(int(n_unsup), n_images)]

    def __len__(self):
        return len(self.samples)

    def sampler(self, indices, batch_size, generator):
        """RepeatSampler(SubsetRandomSampler(indices)) in batches (seg_data.py RepeatSampler, :199-205 of the script): an endless
        stream of index batches, every pass a fresh random permutation drawn from `generator`."""
        indices = np.asarray(indices)

        def gen():
            buf = []
            while True:
                for j in torch.randperm(len(indices), generator=generator).tolist():
                    buf.append(int(indices[j]))
                    if len(buf) == batch_size:
                        yield buf
                        buf = []
        return gen()

    def sup(self, idx):
        return [dict(image_arr=self.samples[i]['image_arr'], labels_arr=self.samples[i]['labels_arr']) for i in idx]

    def unsup(self, idx):
        return [dict(image_arr=self.samples[i]['image_arr'], mask_arr=self.samples[i]['mask_arr']) for i in idx]

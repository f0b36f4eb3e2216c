"""Cases of the device crop / flip / normalise pipeline, shared by oracle/gen_golden.py::gen_input_pipeline
(the reference's own transform classes) and the tests of cutmix_semisup_seg_b200.input_pipeline.DeviceCropFlipNormalize."""
import numpy as np

MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]

CASES = {
    # images larger than the crop, all three flips enabled, labels + mask
    'single_flips': dict(crop_size=(24, 24), crop_offset=None, hflip=True, vflip=True, hvflip=True, mean=MEAN, std=STD, pair=False,
                         seed=11, sizes=[(40, 51), (24, 24), (33, 29), (64, 25), (31, 47), (25, 26)], labels=True, mask=True),
    # images smaller than the crop in one or both dimensions: padding with the alpha-channel standardisation
    'single_padded': dict(crop_size=(32, 40), crop_offset=None, hflip=True, vflip=False, hvflip=False, mean=MEAN, std=STD, pair=False,
                          seed=21, sizes=[(20, 50), (40, 31), (17, 19), (32, 40), (33, 39)], labels=True, mask=True),
    # pairs of crops with an offset (the unsupervised branch), no labels, no standardisation
    'pair_offset': dict(crop_size=(16, 20), crop_offset=(6, 9), hflip=True, vflip=True, hvflip=False, mean=None, std=None, pair=True,
                        seed=31, sizes=[(30, 45), (16, 20), (12, 50), (41, 18)], labels=False, mask=True),
    # pairs with labels, square crop with transposition
    'pair_square': dict(crop_size=(20, 20), crop_offset=(4, 4), hflip=False, vflip=True, hvflip=True, mean=MEAN, std=STD, pair=True,
                        seed=41, sizes=[(28, 36), (19, 33), (50, 21)], labels=True, mask=True),
    # the strong-colour branch of the unsupervised pairs: crop -> flip -> ColorJitter / RandomGrayscale on sample 1 -> normalise
    'pair_colour': dict(crop_size=(24, 28), crop_offset=(3, 5), hflip=True, vflip=False, hvflip=False, mean=MEAN, std=STD, pair=True,
                        seed=51, sizes=[(30, 45), (24, 28), (20, 50), (41, 26), (33, 33), (64, 40), (25, 29), (28, 31), (50, 50), (26, 30)],
                        labels=False, mask=True,
                        colour=dict(brightness=0.4, contrast=0.4, saturation=0.4, hue=0.1, p=0.8, grey_p=0.2), torch_seed=7),
}


def make_samples(case):
    """Seeded uint8 samples: dicts with image_arr (H,W,3) and optionally labels_arr / mask_arr (H,W)."""
    rs = np.random.RandomState(case['seed'] + 1000)
    out = []
    for h, w in case['sizes']:
        s = dict(image_arr=rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8))
        if case['labels']:
            lab = rs.randint(0, 21, size=(h, w)).astype(np.uint8)
            lab[:1] = 255
            s['labels_arr'] = lab
        if case['mask']:
            m = (rs.rand(h, w) > 0.2).astype(np.uint8) * 255
            m[:, :1] = 128
            s['mask_arr'] = m
        out.append(s)
    return out


def reference_statement(samples, params, crop_size, mean, std):
    """numpy statement of the kernel's gather (csrc/input.cu crop_flip_normalize_kernel): pad, crop, flip, normalise per sample."""
    ch, cw = int(crop_size[0]), int(crop_size[1])
    imgs, labs, masks = [], [], []
    for s, p in zip(samples, params):
        img = s['image_arr']
        h0, w0 = img.shape[:2]
        oy, ox = np.mgrid[0:ch, 0:cw]
        cy, cx = oy, ox
        fx, fy, fd = p['flips']
        if fd:
            cy, cx = ox, oy
        if fy:
            cy = ch - 1 - cy
        if fx:
            cx = cw - 1 - cx
        sy, sx = p['pos'][0] + cy - p['pad_top'], p['pos'][1] + cx - p['pad_left']
        inside = (sy >= 0) & (sy < h0) & (sx >= 0) & (sx < w0)
        syc, sxc = np.clip(sy, 0, h0 - 1), np.clip(sx, 0, w0 - 1)
        v = np.where(inside[..., None], np.multiply(img[syc, sxc], 1. / 255, dtype=np.float64), 0.0)
        if mean is not None:
            alpha = np.where(inside, 1.0, 0.0)[..., None] if p['padded'] else 1.0
            v = (v - np.array(mean)[None, None, :] * alpha) / np.array(std)[None, None, :]
        imgs.append(v.transpose(2, 0, 1).astype(np.float32))
        if 'labels_arr' in s:
            labs.append(np.where(inside, s['labels_arr'][syc, sxc], 255)[None].astype(np.int64))
        if 'mask_arr' in s:
            masks.append(np.where(inside, np.multiply(s['mask_arr'][syc, sxc], 1. / 255, dtype=np.float64), 0.0)[None].astype(np.float32))
    out = {'image': np.stack(imgs)}
    if labs:
        out['labels'] = np.stack(labs)
    if masks:
        out['mask'] = np.stack(masks)
    return out

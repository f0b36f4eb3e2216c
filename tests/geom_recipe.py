"""Cases of the device scale / rotation crops (SegCVTransformRandomCropScaleHung, SegCVTransformRandomCropRotateScale),
shared by oracle/gen_golden.py::gen_geom_pipeline (the reference's own transform classes, i.e. cv2.resize / cv2.warpAffine) and the
tests of cutmix_semisup_seg_b200.input_pipeline.DeviceRandomCropScaleHung / DeviceRandomCropRotateScale; plus the numpy statement of
csrc/input.cu's geom_u8_kernel (same integer tables, same integer arithmetic)."""
import numpy as np

MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]

BIG = [(80, 97), (120, 64), (66, 150), (91, 91), (70, 130), (140, 75), (64, 99), (101, 88), (77, 121), (133, 69), (90, 110), (72, 72)]
MIXED = [(80, 97), (30, 64), (66, 25), (20, 22), (70, 130), (40, 41), (64, 99), (33, 88), (77, 29), (133, 69), (18, 110), (72, 72)]

CASES = {
    # --aug_scale_hung, supervised samples: labels + mask, uniform scale, images larger and smaller (padding) than the scaled window
    'hung_single': dict(kind='hung', crop_size=(32, 40), crop_offset=(0, 0), uniform_scale=True, hflip=True, vflip=False, hvflip=False,
                        mean=MEAN, std=STD, pair=False, seed=101, sizes=MIXED + BIG, labels=True, mask=True),
    'hung_single_nonuniform': dict(kind='hung', crop_size=(24, 24), crop_offset=(0, 0), uniform_scale=False, hflip=True, vflip=True,
                                   hvflip=True, mean=None, std=None, pair=False, seed=111, sizes=BIG + MIXED, labels=True, mask=True),
    # unsupervised pairs: second crop scaled, image linear / mask nearest; with and without labels
    'hung_pair': dict(kind='hung', crop_size=(32, 40), crop_offset=(6, 9), uniform_scale=True, hflip=True, vflip=False, hvflip=False,
                      mean=MEAN, std=STD, pair=True, seed=121, sizes=MIXED + BIG, labels=False, mask=True),
    'hung_pair_labels': dict(kind='hung', crop_size=(28, 28), crop_offset=(5, 5), uniform_scale=False, hflip=False, vflip=True,
                             hvflip=True, mean=MEAN, std=STD, pair=True, seed=131, sizes=BIG, labels=True, mask=True),
    # --aug_rot_mag / --aug_max_scale: supervised samples (labels => nearest everywhere), unsupervised singles (random interpolation)
    'rot_single_labels': dict(kind='rot', crop_size=(32, 40), crop_offset=(0, 0), rot_mag=45.0, max_scale=1.3, uniform_scale=True,
                              constrain_rot_scale=True, hflip=True, vflip=True, hvflip=False, mean=MEAN, std=STD, pair=False,
                              seed=141, sizes=MIXED + BIG, labels=True, mask=True),
    'rot_single_nolabels': dict(kind='rot', crop_size=(24, 24), crop_offset=(0, 0), rot_mag=180.0, max_scale=1.5, uniform_scale=False,
                                constrain_rot_scale=True, hflip=True, vflip=True, hvflip=True, mean=None, std=None, pair=False,
                                seed=151, sizes=MIXED + BIG, labels=False, mask=True),
    # unsupervised pairs (linear, reflected image border, zero mask border), constrained and free rotation / scale
    'rot_pair': dict(kind='rot', crop_size=(32, 40), crop_offset=(8, 8), rot_mag=45.0, max_scale=1.1, uniform_scale=True,
                     constrain_rot_scale=True, hflip=True, vflip=True, hvflip=False, mean=MEAN, std=STD, pair=True, seed=161,
                     sizes=MIXED + BIG, labels=False, mask=True),
    'rot_pair_free': dict(kind='rot', crop_size=(28, 28), crop_offset=(4, 7), rot_mag=90.0, max_scale=1.4, uniform_scale=False,
                          constrain_rot_scale=False, hflip=True, vflip=False, hvflip=True, mean=MEAN, std=STD, pair=True, seed=171,
                          sizes=BIG, labels=True, mask=True),
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


def make_transform(case):
    """The device transform of a case, with the reference's generator seeding (crop transform: seed, flip transform: seed + 1)."""
    from cutmix_semisup_seg_b200 import input_pipeline as IP
    common = dict(hflip=case['hflip'], vflip=case['vflip'], hvflip=case['hvflip'], mean=case['mean'], std=case['std'],
                  rng=np.random.RandomState(case['seed']), flip_rng=np.random.RandomState(case['seed'] + 1))
    if case['kind'] == 'hung':
        return IP.DeviceRandomCropScaleHung(case['crop_size'], case['crop_offset'], uniform_scale=case['uniform_scale'], **common)
    return IP.DeviceRandomCropRotateScale(case['crop_size'], case['crop_offset'], case['rot_mag'], case['max_scale'],
                                          uniform_scale=case['uniform_scale'], constrain_rot_scale=case['constrain_rot_scale'], **common)


def drawn(case):
    """(transform, samples, params): parameters drawn sample after sample like the reference's data-set accessor does (pairs: each
    sample appears twice)."""
    tf = make_transform(case)
    samples, params = [], []
    for s in make_samples(case):
        hw = s['image_arr'].shape[:2]
        if case['pair']:
            p0, p1 = tf.draw_pair(hw) if case['kind'] == 'hung' else tf.draw_pair(hw, 'labels_arr' in s)
            samples += [s, s]; params += [p0, p1]
        else:
            params.append(tf.draw_single(hw) if case['kind'] == 'hung' else tf.draw_single(hw, 'labels_arr' in s))
            samples.append(s)
    return tf, samples, params


# ------------------------------------------------------------------------------------------ numpy statement of geom_u8_kernel
def _reflect101(p, n):
    if n == 1:
        return np.zeros_like(p)
    p = p.copy()
    while True:
        bad = (p < 0) | (p >= n)
        if not bad.any():
            return p
        p = np.where(p < 0, -p, np.where(p >= n, 2 * (n - 1) - p, p))


def _sat_short(v):
    return np.clip(v, -32768, 32767)


def geom_u8(sample, p, crop_size):
    """One sample through the kernel's arithmetic (before the flips): RGBA uint8 (h, w, 4), labels uint8 | None, mask uint8 | None."""
    from cutmix_semisup_seg_b200.input_pipeline import resize_tables, warp_tables, NEAREST, LINEAR
    H, W = int(crop_size[0]), int(crop_size[1])
    img = sample['image_arr'].astype(np.int64)
    h0, w0 = img.shape[:2]
    lab, msk = sample.get('labels_arr'), sample.get('mask_arr')
    ry, rx = np.mgrid[0:H, 0:W]
    if p['mode'] == 0:
        (xn, xl, xa, yn, yl, yb), area2 = resize_tables(p['src_size'], (H, W))
        sh, sw = p['src_size']

        def window(plane, y, x, outside):                      # pixel (y, x) of the window of the virtually padded image
            sy, sx = p['pos'][0] + y - p['pad_top'], p['pos'][1] + x - p['pad_left']
            ok = (sy >= 0) & (sy < h0) & (sx >= 0) & (sx < w0)
            v = plane[np.clip(sy, 0, h0 - 1), np.clip(sx, 0, w0 - 1)]
            return np.where(ok[..., None] if v.ndim == 3 else ok, v, outside)
        rgba_src = np.concatenate([img, np.full((h0, w0, 1), 255, np.int64)], axis=2)
        a0, a1 = (xa & 0xffff).astype(np.int64)[None, :], ((xa >> 16) & 0xffff).astype(np.int64)[None, :]
        b0, b1 = (yb & 0xffff).astype(np.int64)[:, None], ((yb >> 16) & 0xffff).astype(np.int64)[:, None]
        x0 = np.broadcast_to(xl.astype(np.int64)[None, :], (H, W)); x1 = np.minimum(x0 + 1, sw - 1)
        y0 = np.broadcast_to(np.clip(yl.astype(np.int64), 0, sh - 1)[:, None], (H, W))
        y1 = np.broadcast_to(np.clip(yl.astype(np.int64) + 1, 0, sh - 1)[:, None], (H, W))
        ynn = np.broadcast_to(yn.astype(np.int64)[:, None], (H, W)); xnn = np.broadcast_to(xn.astype(np.int64)[None, :], (H, W))

        def resample(plane, interp, outside):
            ex = (lambda a: a[..., None]) if plane.ndim == 3 else (lambda a: a)
            if interp == NEAREST:
                return window(plane, ynn, xnn, outside)
            if interp == LINEAR and not area2:
                r0 = window(plane, y0, x0, outside) * ex(a0) + window(plane, y0, x1, outside) * ex(a1)
                r1 = window(plane, y1, x0, outside) * ex(a0) + window(plane, y1, x1, outside) * ex(a1)
                return (((ex(b0) * (r0 >> 4)) >> 16) + ((ex(b1) * (r1 >> 4)) >> 16) + 2) >> 2
            return (window(plane, 2 * ry, 2 * rx, outside) + window(plane, 2 * ry, 2 * rx + 1, outside) +
                    window(plane, 2 * ry + 1, 2 * rx, outside) + window(plane, 2 * ry + 1, 2 * rx + 1, outside) + 2) >> 2
        rgba = resample(rgba_src, p['image_interp'], 0)
        out_lab = None if lab is None else resample(lab.astype(np.int64), NEAREST, 255)
        out_msk = None if msk is None else resample(msk.astype(np.int64), p['mask_interp'], 0)
    else:
        ad, bd, X0, Y0 = (t.astype(np.int64) for t in warp_tables(p['matrix'], (H, W)))
        nx = _sat_short((X0[:, None] + 512 + ad[None, :]) >> 10); ny = _sat_short((Y0[:, None] + 512 + bd[None, :]) >> 10)
        n_in = (nx >= 0) & (nx < w0) & (ny >= 0) & (ny < h0)
        LX = (X0[:, None] + 16 + ad[None, :]) >> 5; LY = (Y0[:, None] + 16 + bd[None, :]) >> 5
        sx, sy = _sat_short(LX >> 5), _sat_short(LY >> 5)
        fx, fy = LX & 31, LY & 31
        w = [(32 - fy) * (32 - fx) * 32, (32 - fy) * fx * 32, fy * (32 - fx) * 32, fy * fx * 32]
        if p['image_interp'] == NEAREST:
            rgb = img[_reflect101(ny, h0), _reflect101(nx, w0)]
        else:
            xa_, xb_ = _reflect101(sx, w0), _reflect101(sx + 1, w0)
            ya_, yb_ = _reflect101(sy, h0), _reflect101(sy + 1, h0)
            taps = [img[ya_, xa_], img[ya_, xb_], img[yb_, xa_], img[yb_, xb_]]
            rgb = (sum(t * wi[..., None] for t, wi in zip(taps, w)) + (1 << 14)) >> 15
        rgba = np.concatenate([rgb, np.full((H, W, 1), 255, np.int64)], axis=2)

        def const_nearest(plane, cval):
            return np.where(n_in, plane[np.clip(ny, 0, h0 - 1), np.clip(nx, 0, w0 - 1)], cval)
        out_lab = None if lab is None else const_nearest(lab.astype(np.int64), 255)
        if msk is None:
            out_msk = None
        elif p['mask_interp'] == NEAREST:
            out_msk = const_nearest(msk.astype(np.int64), 0)
        else:
            m = msk.astype(np.int64)

            def tap(yy, xx):
                ok = (xx >= 0) & (xx < w0) & (yy >= 0) & (yy < h0)
                return np.where(ok, m[np.clip(yy, 0, h0 - 1), np.clip(xx, 0, w0 - 1)], 0)
            acc = tap(sy, sx) * w[0] + tap(sy, sx + 1) * w[1] + tap(sy + 1, sx) * w[2] + tap(sy + 1, sx + 1) * w[3]
            all_out = (sx >= w0) | (sx + 1 < 0) | (sy >= h0) | (sy + 1 < 0)
            out_msk = np.where(all_out, 0, (acc + (1 << 14)) >> 15)
    u8 = lambda a: None if a is None else a.astype(np.uint8)
    return u8(rgba), u8(out_lab), u8(out_msk)


def flip(a, flags):
    """SegCVTransformRandomFlip.flip_image (seg_transforms_cv.py:467-474)."""
    if flags[0]:
        a = a[:, ::-1]
    if flags[1]:
        a = a[::-1, ...]
    if flags[2]:
        a = np.swapaxes(a, 0, 1)
    return a


def statement(samples, params, crop_size, mean, std):
    """The whole device chain in numpy: geom_u8 -> flip -> normalise-to-tensor (float64 arithmetic, one rounding)."""
    imgs, labs, masks = [], [], []
    for s, p in zip(samples, params):
        rgba, lab, msk = geom_u8(s, p, crop_size)
        rgba = flip(rgba, p['flips'])
        v = np.multiply(rgba[..., :3], 1. / 255, dtype=np.float64)
        if mean is not None:
            alpha = np.multiply(rgba[..., 3:4], 1. / 255, dtype=np.float64)
            v = (v - np.array(mean)[None, None, :] * alpha) / np.array(std)[None, None, :]
        imgs.append(v.transpose(2, 0, 1).astype(np.float32))
        if lab is not None:
            labs.append(flip(lab, p['flips'])[None].astype(np.int64))
        if msk is not None:
            masks.append(np.multiply(flip(msk, p['flips']), 1. / 255, dtype=np.float64)[None].astype(np.float32))
    out = {'image': np.stack(imgs)}
    if labs:
        out['labels'] = np.stack(labs)
    if masks:
        out['mask'] = np.stack(masks)
    return out

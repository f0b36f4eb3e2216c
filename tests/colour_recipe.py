"""numpy statement of the device colour-jitter kernels (csrc/input.cu: colour_apply_kernel / colour_lum_sum_kernel): the uint8
arithmetic of the reference's strong colour augmentation, i.e. torchvision's ColorJitter + RandomGrayscale on PIL images
(train_seg_semisup_mask_mt.py:169-179 -> datapipe/seg_transforms_cv.py:541-586 SegCVTransformTVT), which bottoms out in Pillow's C
code.  Pillow's sources are not in the image, so the algorithms below restate its published ones (libImaging/Blend.c ImagingBlend,
Convert.c rgb2l / rgb2hsv_row / hsv2rgb, ImageEnhance.py, ImageStat.py) INCLUDING their float / double mixing, and
tests/test_colour_jitter.py pins them against the installed Pillow 12.2 / torchvision 0.26 -- exhaustively over all 2^24 colours
for the two HSV conversions."""
import numpy as np

f32, f64 = np.float32, np.float64
BRIGHTNESS, CONTRAST, SATURATION, HUE = 0, 1, 2, 3


def lum(rgb):
    """Image.convert('L'): ITU-R 601-2 luma in 16.16 fixed point."""
    r, g, b = (rgb[..., i].astype(np.int64) for i in range(3))
    return ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8)


def blend(degenerate, image, alpha):
    """Image.blend(degenerate, image, alpha): in1 + alpha * (in2 - in1) in C float; truncation inside [0, 1], clipping outside."""
    a = f32(alpha)
    d, x = degenerate.astype(np.int64), image.astype(np.int64)
    t = (d.astype(f32) + (a * (x - d).astype(f32)).astype(f32)).astype(f32)
    if 0.0 <= float(a) <= 1.0:
        return t.astype(np.int64).astype(np.uint8)
    return np.where(t <= 0, 0, np.where(t >= 255, 255, t.astype(np.int64))).astype(np.uint8)


def contrast_mean(rgb):
    """ImageEnhance.Contrast: int(ImageStat.Stat(image.convert('L')).mean[0] + 0.5)."""
    L = lum(rgb)
    return int(float(int(L.astype(np.int64).sum())) / L.size + 0.5)


def rgb_to_hsv(rgb):
    r, g, b = (rgb[..., i].astype(np.int64) for i in range(3))
    maxc, minc = np.maximum(np.maximum(r, g), b), np.minimum(np.minimum(r, g), b)
    with np.errstate(divide='ignore', invalid='ignore'):
        cr = (maxc - minc).astype(f32)
        s = cr / maxc.astype(f32)
        rc, gc, bc = ((maxc - c).astype(f32) / cr for c in (r, g, b))
        h = np.where(r == maxc, (bc - gc).astype(f32),
                     np.where(g == maxc, (2.0 + rc.astype(f64) - bc.astype(f64)).astype(f32),
                              (4.0 + gc.astype(f64) - rc.astype(f64)).astype(f32)))
        h = np.fmod(h.astype(f64) / 6.0 + 1.0, 1.0).astype(f32)
        uh = np.clip((h.astype(f64) * 255.0).astype(np.int64), 0, 255)
        us = np.clip((s.astype(f64) * 255.0).astype(np.int64), 0, 255)
    grey = maxc == minc
    return np.stack([np.where(grey, 0, uh), np.where(grey, 0, us), maxc], axis=-1).astype(np.uint8)


def _c_round(x):
    return np.where(x >= 0, np.floor(x + 0.5), np.ceil(x - 0.5)).astype(np.int64)


def hsv_to_rgb(hsv):
    h, s, v = hsv[..., 0], hsv[..., 1], hsv[..., 2]
    h6 = h.astype(f32).astype(f64) * 6.0 / 255.0
    i = np.floor(h6).astype(np.int64)
    f = (h6 - i.astype(f32).astype(f64)).astype(f32)
    fs = (s.astype(f32).astype(f64) / 255.0).astype(f32)
    vf = v.astype(f32).astype(f64)
    fsf = (fs * f).astype(f32).astype(f64)                                   # float * float stays float in C
    p = np.clip(_c_round(vf * (1.0 - fs.astype(f64))), 0, 255)
    q = np.clip(_c_round(vf * (1.0 - fsf)), 0, 255)
    # fs * (1.0 - f): (1.0 - f) is double, so the product is double
    t = np.clip(_c_round(vf * (1.0 - fs.astype(f64) * (1.0 - f.astype(f64)))), 0, 255)
    vi = v.astype(np.int64)
    m = i % 6
    r = np.choose(m, [vi, q, p, p, t, vi]); g = np.choose(m, [t, vi, vi, q, p, p]); b = np.choose(m, [p, p, t, vi, vi, q])
    zero = s == 0
    return np.stack([np.where(zero, vi, r), np.where(zero, vi, g), np.where(zero, vi, b)], axis=-1).astype(np.uint8)


def hue_shift(hue_factor):
    """torchvision _functional_pil.adjust_hue: np.int32(hue_factor * 255).astype(np.uint8), added to H with uint8 wrap-around."""
    return int(np.int32(hue_factor * 255).astype(np.uint8))


def apply(rgb, params):
    """One image (H, W, 3) uint8 through the drawn parameters of DeviceColourJitter.draw (ops in the drawn order, then grey)."""
    out = rgb.copy()
    for op, fac in params['ops']:
        if op == BRIGHTNESS:
            out = blend(np.zeros_like(out), out, fac)
        elif op == CONTRAST:
            out = blend(np.full_like(out, contrast_mean(out)), out, fac)
        elif op == SATURATION:
            out = blend(np.repeat(lum(out)[..., None], 3, axis=2), out, fac)
        elif op == HUE:
            hsv = rgb_to_hsv(out)
            hsv[..., 0] = (hsv[..., 0].astype(np.int64) + hue_shift(fac)).astype(np.uint8)
            out = hsv_to_rgb(hsv)
    if params['grey']:
        out = np.repeat(lum(out)[..., None], 3, axis=2)
    return out

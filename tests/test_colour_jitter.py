"""CPU: the colour-jitter algorithm (tests/colour_recipe.py = the numpy statement of csrc/input.cu's colour kernels) against the
installed Pillow / torchvision, which is what the reference's strong colour augmentation executes (SegCVTransformTVT,
datapipe/seg_transforms_cv.py:541-586).  The two HSV conversions are compared on ALL 2^24 colours."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(__file__)
sys.path.insert(0, HERE)
import colour_recipe as CR  # noqa: E402


def _all_colours():
    a = np.arange(256, dtype=np.uint8)
    x, y, z = np.meshgrid(a, a, a, indexing='ij')
    return np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1).reshape(4096, 4096, 3)


def test_hsv_conversions_match_pillow_on_all_colours():
    from PIL import Image
    cube = _all_colours()
    assert np.array_equal(CR.rgb_to_hsv(cube), np.array(Image.fromarray(cube).convert('HSV')))
    assert np.array_equal(CR.hsv_to_rgb(cube), np.array(Image.fromarray(cube, 'HSV').convert('RGB')))
    assert np.array_equal(CR.lum(cube), np.array(Image.fromarray(cube).convert('L')))


@pytest.mark.parametrize('factor', [0.0, 0.6, 0.83, 1.0, 1.17, 1.4, 2.5])
def test_enhancers_match_torchvision_functional_on_pil_images(factor):
    import torchvision.transforms.functional as TF
    from PIL import Image
    rs = np.random.RandomState(int(factor * 100))
    img = rs.randint(0, 256, size=(61, 83, 3)).astype(np.uint8)
    img[:3, :3] = 0; img[3:6, :3] = 255
    pil = Image.fromarray(img)
    assert np.array_equal(CR.blend(np.zeros_like(img), img, factor), np.array(TF.adjust_brightness(pil, factor)))
    assert np.array_equal(CR.blend(np.full_like(img, CR.contrast_mean(img)), img, factor), np.array(TF.adjust_contrast(pil, factor)))
    assert np.array_equal(CR.blend(np.repeat(CR.lum(img)[..., None], 3, 2), img, factor), np.array(TF.adjust_saturation(pil, factor)))
    hue = (factor - 1.0) / 3.0                       # a value in [-0.5, 0.5]
    got = CR.apply(img, dict(ops=[(CR.HUE, hue)], grey=False))
    assert np.array_equal(got, np.array(TF.adjust_hue(pil, hue)))
    assert np.array_equal(CR.apply(img, dict(ops=[], grey=True)), np.array(TF.rgb_to_grayscale(pil, 3)))


def test_parameter_draws_follow_torchvision_and_the_table_layout_matches_the_header():
    """DeviceColourJitter.draw consumes torch's generator like tvt.Compose([RandomApply([ColorJitter]), RandomGrayscale]) does:
    applying the drawn parameters with the numpy statement == applying the torchvision transforms with the same seed."""
    import torchvision.transforms as tvt
    from PIL import Image
    from cutmix_semisup_seg_b200.input_pipeline import DeviceColourJitter
    rs = np.random.RandomState(3)
    imgs = [rs.randint(0, 256, size=(17, 23, 3)).astype(np.uint8) for _ in range(40)]
    xf = tvt.Compose([tvt.RandomApply([tvt.ColorJitter(0.4, 0.4, 0.4, 0.1)], p=0.8), tvt.RandomGrayscale(p=0.2)])
    torch.manual_seed(123)
    want = [np.array(xf(Image.fromarray(im))) for im in imgs]
    cj = DeviceColourJitter(0.4, 0.4, 0.4, 0.1, 0.8, 0.2)
    torch.manual_seed(123)
    params = [cj.draw() for _ in imgs]
    for im, p, w in zip(imgs, params, want):
        assert np.array_equal(CR.apply(im, p), w)
    assert any(p['grey'] for p in params) and any(not p['ops'] for p in params) and any(len(p['ops']) == 4 for p in params)
    tab = DeviceColourJitter.table(params)
    assert tab.dtype.itemsize == 56 and tab.dtype.fields['factor'][1] == 20 and tab.dtype.fields['hue_shift'][1] == 36 and \
        tab.dtype.fields['grey'][1] == 52
    # a disabled factor (magnitude 0) draws nothing, like torchvision's `None` ranges
    assert DeviceColourJitter(0.4, 0.0, 0.4, 0.0).ranges[1] is None and DeviceColourJitter(0.4, 0.0, 0.4, 0.0).ranges[3] is None

"""-m gpu: csrc/input.cu (device-side SegCVTransformNormalizeToTensor) bit-exact against the reference's
numpy arithmetic restated in tests/test_input_pipeline.py."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
from test_input_pipeline import MEAN, STD, check_backend, make_batch, reference_transform_single  # noqa: E402


@pytest.mark.gpu
def test_cuda_kernels_are_bit_exact_with_the_reference_arithmetic():
    from cutmix_semisup_seg_b200 import ops, input_pipeline
    dev = torch.device('cuda:0')
    check_backend(ops.default_backend(), to_dev=lambda t: t.to(dev))
    img, lab, mask = make_batch(2, 17, 19, 3, seed=1)
    tf = input_pipeline.DeviceNormalizeToTensor(MEAN, STD)
    out = tf(dict(image_arr=torch.from_numpy(img).pin_memory(), labels_arr=torch.from_numpy(lab).to(dev),
                  mask_arr=torch.from_numpy(mask).to(dev), index=torch.arange(2)))
    assert set(out) == {'image', 'labels', 'mask', 'index'}
    want = reference_transform_single(dict(image_arr=img[1], labels_arr=lab[1], mask_arr=mask[1]), MEAN, STD)
    assert np.array_equal(out['image'][1].cpu().numpy(), want['image'])
    assert np.array_equal(out['labels'][1].cpu().numpy(), want['labels'])
    assert np.array_equal(out['mask'][1].cpu().numpy(), want['mask'])


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['single_flips', 'single_padded', 'pair_offset', 'pair_square'])
def test_crop_flip_normalize_kernel_matches_the_reference_transform_classes(name):
    """b2_crop_flip_normalize (one gather pass over variable-size uint8 images on the device) against the outputs of the
    reference's own SegCVTransformRandomCrop -> RandomFlip -> NormalizeToTensor classes (tests/golden/input_pipeline.npz)."""
    import input_recipe as IR
    from test_input_pipeline import _drawn
    dev = torch.device('cuda:0')
    gold = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'input_pipeline.npz'))
    case = IR.CASES[name]
    tf, samples, params = _drawn(case)
    dev_samples = [{k: torch.from_numpy(v).to(dev) for k, v in s.items()} for s in samples]
    out = tf(dev_samples, params)
    assert np.array_equal(out['image'].cpu().numpy(), gold[name + '.image'])
    if case['labels']:
        assert out['labels'].dtype == torch.int64 and np.array_equal(out['labels'].cpu().numpy(), gold[name + '.labels'])
    else:
        assert 'labels' not in out
    if case['mask']:
        assert np.array_equal(out['mask'].cpu().numpy(), gold[name + '.mask'])
    # host (pinned) samples are copied first: same result
    out2 = tf([{k: torch.from_numpy(v).pin_memory() for k, v in s.items()} for s in samples], params)
    assert torch.equal(out2['image'], out['image'])


@pytest.mark.gpu
def test_strong_colour_branch_matches_the_reference_transform_chain():
    """crop -> flip -> ColorJitter / RandomGrayscale -> normalise on the device (b2_crop_flip_u8, b2_colour_jitter,
    b2_normalize_to_tensor) against the reference's SegCVTransformRandomCrop -> RandomFlip -> SegCVTransformTVT(torchvision) ->
    NormalizeToTensor chain on the same seeds (tests/golden/input_pipeline.npz, case pair_colour)."""
    import input_recipe as IR
    from test_input_pipeline import _drawn
    dev = torch.device('cuda:0')
    gold = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'input_pipeline.npz'))
    case = IR.CASES['pair_colour']
    tf, samples, params = _drawn(case)
    dev_samples = [{k: torch.from_numpy(v).to(dev) for k, v in s.items()} for s in samples]
    out = tf(dev_samples, params, colour=tf.colour, colour_params=tf.cparams)
    assert np.array_equal(out['image'].cpu().numpy(), gold['pair_colour.image'])
    assert np.array_equal(out['mask'].cpu().numpy(), gold['pair_colour.mask'])


@pytest.mark.gpu
def test_colour_kernels_match_the_pillow_arithmetic_on_all_colours():
    """Every operation of b2_colour_jitter on all 2^24 colours (a 4096 x 4096 image) against the numpy statement that
    tests/test_colour_jitter.py pins to Pillow: hue shifts (RGB -> HSV -> RGB), brightness / saturation inside and outside
    [0, 1], contrast (image-wide mean), grey, and a four-operation chain."""
    import colour_recipe as CR
    from cutmix_semisup_seg_b200.input_pipeline import DeviceColourJitter
    dev = torch.device('cuda:0')
    a = np.arange(256, dtype=np.uint8)
    x, y, z = np.meshgrid(a, a, a, indexing='ij')
    cube = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1).reshape(4096, 4096, 3)
    cj = DeviceColourJitter()
    cases = [dict(ops=[(CR.HUE, 0.0)], grey=False), dict(ops=[(CR.HUE, 0.1)], grey=False), dict(ops=[(CR.HUE, -0.37)], grey=False),
             dict(ops=[(CR.BRIGHTNESS, 0.6)], grey=False), dict(ops=[(CR.BRIGHTNESS, 1.4)], grey=False),
             dict(ops=[(CR.SATURATION, 0.77)], grey=False), dict(ops=[(CR.SATURATION, 1.31)], grey=True),
             dict(ops=[(CR.CONTRAST, 0.61)], grey=False), dict(ops=[(CR.CONTRAST, 1.39)], grey=False), dict(ops=[], grey=True),
             dict(ops=[(CR.SATURATION, 1.2), (CR.HUE, 0.05), (CR.CONTRAST, 0.8), (CR.BRIGHTNESS, 1.1)], grey=False)]
    for start in range(0, len(cases), 4):
        chunk = cases[start:start + 4]
        img = torch.from_numpy(np.stack([cube] * len(chunk))).to(dev)
        cj(img, chunk)
        got = img.cpu().numpy()
        for i, p in enumerate(chunk):
            assert np.array_equal(got[i], CR.apply(cube, p)), p
    # RGBA pixels: the alpha plane is left alone
    rgba = torch.from_numpy(np.concatenate([cube[:64], np.full((64, 4096, 1), 200, np.uint8)], axis=2)[None]).to(dev).contiguous()
    cj(rgba, [cases[-1]])
    assert np.array_equal(rgba.cpu().numpy()[0, ..., :3], CR.apply(cube[:64], cases[-1])) and int(rgba[..., 3].min()) == 200


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['hung_single', 'hung_single_nonuniform', 'hung_pair', 'hung_pair_labels', 'rot_single_labels',
                                  'rot_single_nolabels', 'rot_pair', 'rot_pair_free'])
def test_scale_rotation_crop_kernel_matches_the_reference_transform_classes(name):
    """b2_geom_u8 (cv2.resize / cv2.warpAffine in OpenCV's fixed-point arithmetic + flips, one gather over variable-size uint8 images)
    -> b2_normalize_to_tensor against the outputs of the reference's own SegCVTransformRandomCropScaleHung /
    SegCVTransformRandomCropRotateScale -> RandomFlip -> NormalizeToTensor classes (tests/golden/geom_pipeline.npz), and the raw
    RGBA / label / mask bytes against the numpy statement of the kernel."""
    import geom_recipe as GR
    dev = torch.device('cuda:0')
    gold = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'geom_pipeline.npz'))
    case = GR.CASES[name]
    tf, samples, params = GR.drawn(case)
    dev_samples = [{k: torch.from_numpy(v).to(dev) for k, v in s.items()} for s in samples]
    out = tf(dev_samples, params)
    assert np.array_equal(out['image'].cpu().numpy(), gold[name + '.image'])
    if case['labels']:
        assert out['labels'].dtype == torch.int64 and np.array_equal(out['labels'].cpu().numpy(), gold[name + '.labels'])
    else:
        assert 'labels' not in out
    if case['mask']:
        assert np.array_equal(out['mask'].cpu().numpy(), gold[name + '.mask'])
    # the un-normalised bytes of the kernel itself
    ent, tab = tf.tables(dev_samples, params)
    from cutmix_semisup_seg_b200 import ops
    h, w = case['crop_size']
    rgba, _, _ = ops.default_backend().geom_u8(torch.from_numpy(ent.view('u1').copy()).to(dev), torch.from_numpy(tab).to(dev),
                                               len(samples), h, w, False, False, dev)
    rgba = rgba.cpu().numpy()
    for i, (s, p) in enumerate(zip(samples, params)):
        assert np.array_equal(rgba[i], GR.flip(GR.geom_u8(s, p, case['crop_size'])[0], p['flips'])), i
    # host (pinned) samples are copied first: same result
    out2 = tf([{k: torch.from_numpy(v).pin_memory() for k, v in s.items()} for s in samples], params)
    assert torch.equal(out2['image'], out['image'])


@pytest.mark.gpu
def test_scale_crop_full_size_against_cv2_and_colour_chain():
    """Pascal-recipe sizes (321 x 321 crops of ~500 x 375 images, every scale factor 0.5 ... 1.5): the kernel's RGBA bytes against
    cv2.resize itself; then the colour-jitter branch on top of the scaled crops runs through the same kernels as the plain-crop
    pipeline (bytes = numpy statement of geom + tests/colour_recipe.py)."""
    cv2 = pytest.importorskip('cv2')
    cv2.setNumThreads(0)
    import colour_recipe as CR
    import geom_recipe as GR
    from cutmix_semisup_seg_b200 import input_pipeline as IP, ops
    dev = torch.device('cuda:0')
    rng = np.random.RandomState(5)
    crop = (321, 321)
    tf = IP.DeviceRandomCropScaleHung(crop, (0, 0), hflip=True, rng=np.random.RandomState(1), flip_rng=np.random.RandomState(2))
    samples, params = [], []
    for f10 in range(5, 16):
        sc = int(np.round(321 / (f10 / 10.0)))
        h0, w0 = sc + int(rng.randint(0, 40)), sc + int(rng.randint(0, 60))
        s = dict(image_arr=rng.randint(0, 256, size=(h0, w0, 3)).astype(np.uint8), mask_arr=rng.randint(0, 256, size=(h0, w0)).astype(np.uint8))
        pos = (int(rng.randint(0, h0 - sc + 1)), int(rng.randint(0, w0 - sc + 1)))
        samples.append(s)
        params.append(dict(mode=0, pad_top=0, pad_left=0, padded=0, pos=pos, src_size=(sc, sc), image_interp=IP.LINEAR,
                           mask_interp=IP.LINEAR, flips=(False, False, False)))
    dev_samples = [{k: torch.from_numpy(v).to(dev) for k, v in s.items()} for s in samples]
    ent, tab = tf.tables(dev_samples, params)
    be = ops.default_backend()
    rgba, _, mask = be.geom_u8(torch.from_numpy(ent.view('u1').copy()).to(dev), torch.from_numpy(tab).to(dev), len(samples), 321, 321,
                               False, True, dev)
    got, got_m = rgba.cpu().numpy(), mask.cpu().numpy()
    for i, (s, p) in enumerate(zip(samples, params)):
        (y, x), (sh, sw) = p['pos'], p['src_size']
        want = cv2.resize(s['image_arr'][y:y + sh, x:x + sw], (321, 321), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(got[i, ..., :3], want), i
        want_m = cv2.resize(s['mask_arr'][y:y + sh, x:x + sw], (321, 321), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(got_m[i, 0], np.multiply(want_m, 1. / 255, dtype=np.float64).astype(np.float32)), i
    # colour jitter on the scaled crops (strong-colour branch of the Pascal recipe)
    cj = IP.DeviceColourJitter()
    torch.manual_seed(3)
    cparams = [cj.draw() for _ in samples]
    tf.mean, tf.std = GR.MEAN, GR.STD
    out = tf(dev_samples, params, colour=cj, colour_params=cparams)
    for i, (s, p, cp) in enumerate(zip(samples, params, cparams)):
        u8 = GR.geom_u8(s, p, crop)[0]
        j = CR.apply(u8[..., :3], cp)
        v = (np.multiply(j, 1. / 255, dtype=np.float64) - np.array(GR.MEAN)[None, None, :]) / np.array(GR.STD)[None, None, :]
        assert np.array_equal(out['image'][i].cpu().numpy(), v.transpose(2, 0, 1).astype(np.float32)), i

"""CPU: the data-format boundary.  The reference's SegCVTransformNormalizeToTensor
(datapipe/seg_transforms_cv.py:587-672) cannot be imported here (it needs scikit-image, pinned 0.16.2 in environment.yml),
so its arithmetic is restated below line by line with numpy -- img_as_float for uint8 is `np.multiply(image, 1. / 255,
dtype=float64)` (skimage/util/dtype.py, `convert`) -- and the kernels' algorithm (tests/_emu_backend.py states it a second
time) must agree with it bit for bit; the `-m gpu` half runs the same comparison on the CUDA kernels.  parity unpinned for
the scikit-image call (the package is absent); everything else follows the reference's own lines."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(__file__)
sys.path.insert(0, HERE)
MEAN = np.array([0.485, 0.456, 0.406])          # the networks' MEAN / STD (deeplab2.py, deeplab3plus.py)
STD = np.array([0.229, 0.224, 0.225])


def img_as_float(a):
    return np.multiply(a, 1. / 255, dtype=np.float64)


def reference_transform_single(sample, mean, std):
    """seg_transforms_cv.py:592-623, verbatim arithmetic."""
    sample = sample.copy()
    image = img_as_float(sample['image_arr'])                                                   # :596
    if image.shape[2] == 4:
        alpha_channel = image[:, :, 3:4]                                                        # :601
        image = image[:, :, :3]
        if mean is not None and std is not None:
            image = (image - (mean[None, None, :] * alpha_channel)) / std[None, None, :]        # :606
    else:
        if mean is not None and std is not None:
            image = (image - mean[None, None, :]) / std[None, None, :]                          # :610
    assert image.shape[2] == 3
    sample['image'] = image.transpose(2, 0, 1).astype(np.float32)                               # :614
    del sample['image_arr']
    if 'labels_arr' in sample:
        sample['labels'] = sample['labels_arr'][None, ...].astype(np.int64)                     # :617
        del sample['labels_arr']
    if 'mask_arr' in sample:
        sample['mask'] = img_as_float(sample['mask_arr'])[None, ...].astype(np.float32)         # :620
        del sample['mask_arr']
    return sample


def make_batch(n, h, w, cin, seed):
    rs = np.random.RandomState(seed)
    img = rs.randint(0, 256, size=(n, h, w, cin)).astype(np.uint8)
    edge = np.array([[0] * cin, [255] * cin, [1] * cin, [254] * cin], dtype=np.uint8)
    img.reshape(-1, cin)[:min(4, n * h * w)] = edge[:min(4, n * h * w)]
    lab = rs.randint(0, 21, size=(n, h, w)).astype(np.uint8); lab[:, :2] = 255
    mask = (rs.rand(n, h, w) > 0.2).astype(np.uint8) * 255; mask[:, :, 0] = 128
    return img, lab, mask


def check_backend(be, to_dev=lambda t: t):
    for cin, h, w, norm in ((3, 9, 13, True), (4, 6, 5, True), (3, 33, 47, False), (3, 1, 1, True)):
        img, lab, mask = make_batch(3, h, w, cin, seed=h * w + cin)
        mean, std = (MEAN, STD) if norm else (None, None)
        want = [reference_transform_single(dict(image_arr=img[i], labels_arr=lab[i], mask_arr=mask[i]), mean, std)
                for i in range(len(img))]
        got_img = be.normalize_to_tensor(to_dev(torch.from_numpy(img)), mean, std).cpu().numpy()
        got_lab = be.labels_to_tensor(to_dev(torch.from_numpy(lab))).cpu().numpy()
        got_mask = be.mask_to_tensor(to_dev(torch.from_numpy(mask))).cpu().numpy()
        assert got_img.dtype == np.float32 and got_lab.dtype == np.int64 and got_mask.dtype == np.float32
        assert np.array_equal(got_img, np.stack([s['image'] for s in want]))                   # bit-exact
        assert np.array_equal(got_lab, np.stack([s['labels'] for s in want]))
        assert np.array_equal(got_mask, np.stack([s['mask'] for s in want]))


def test_kernel_algorithm_is_bit_exact_with_the_reference_arithmetic():
    from _emu_backend import EmuBackend
    check_backend(EmuBackend())


def test_wrong_channel_count_is_rejected_like_the_reference():
    from cutmix_semisup_seg_b200 import lib as L
    with pytest.raises(L.B2Error, match='should have 3 channels'):                       # seg_transforms_cv.py:653-656
        L.call('b2_normalize_to_tensor', 0x1000, 1, 4, 4, 2, None, None, 0x1000, None)
    with pytest.raises(L.B2Error, match='together'):
        L.call('b2_normalize_to_tensor', 0x1000, 1, 4, 4, 3, (L.ctypes.c_double * 3)(0, 0, 0), None, 0x1000, None)


# ------------------------------------------------------------------------------------------ crop + flip + normalise
def _drawn(case):
    """Parameters drawn by DeviceCropFlipNormalize in the reference's order, and the samples (pairs: each sample twice)."""
    import input_recipe as IR
    from cutmix_semisup_seg_b200.input_pipeline import DeviceCropFlipNormalize
    tf = DeviceCropFlipNormalize(case['crop_size'], case['crop_offset'], case['hflip'], case['vflip'], case['hvflip'], case['mean'],
                                 case['std'], crop_rng=np.random.RandomState(case['seed']),
                                 flip_rng=np.random.RandomState(case['seed'] + 1))
    samples, params = [], []
    colour, cparams = None, []
    if case.get('colour'):
        from cutmix_semisup_seg_b200.input_pipeline import DeviceColourJitter
        colour = DeviceColourJitter(**case['colour'])
        torch.manual_seed(case['torch_seed'])
    tf.colour, tf.cparams = colour, cparams
    for s in IR.make_samples(case):
        # NOTE the reference draws crop parameters of ALL samples from the crop transform's generator and flips from the flip
        # transform's, sample after sample (transform_single / transform_pair are called per sample by the data set accessor)
        if case['pair']:
            p0, p1 = tf.draw_pair(s['image_arr'].shape[:2])
            samples += [s, s]; params += [p0, p1]
            if colour is not None:                  # SegCVTransformTVT(apply_pair0=False, apply_pair1=True)
                cparams += [dict(ops=[], grey=False), colour.draw()]
        else:
            samples.append(s); params.append(tf.draw_single(s['image_arr'].shape[:2]))
    return tf, samples, params


def _statement(case, samples, params, cparams):
    """numpy statement of the device pipeline: crop / flip gather, optional colour jitter on the uint8 crop, normalise."""
    import input_recipe as IR
    import colour_recipe as CR
    if not case.get('colour'):
        return IR.reference_statement(samples, params, case['crop_size'], case['mean'], case['std'])
    raw = IR.reference_statement(samples, params, case['crop_size'], None, None)            # crops as [0,1] floats = u8 / 255
    u8 = np.rint(raw['image'].transpose(0, 2, 3, 1).astype(np.float64) * 255.0).astype(np.uint8)
    alpha = []
    for s, p in zip(samples, params):                      # alpha plane of padded samples: where the gather left the source
        probe = dict(image_arr=np.full(s['image_arr'].shape, 255, np.uint8))
        a = IR.reference_statement([probe], [p], case['crop_size'], None, None)['image'][0, 0]
        alpha.append(a)
    out = []
    for img, cp, a in zip(u8, cparams, alpha):
        j = CR.apply(img, cp)
        v = np.multiply(j, 1. / 255, dtype=np.float64)
        v = (v - np.array(case['mean'])[None, None, :] * a.astype(np.float64)[..., None]) / np.array(case['std'])[None, None, :]
        out.append(v.transpose(2, 0, 1).astype(np.float32))
    raw['image'] = np.stack(out)
    return raw


@pytest.mark.parametrize('name', ['single_flips', 'single_padded', 'pair_offset', 'pair_square', 'pair_colour'])
def test_crop_flip_normalize_algorithm_matches_the_reference_transform_classes(name):
    """The host-side parameter draws + the kernel's gather (stated in numpy, tests/input_recipe.py) reproduce, bit for bit,
    what the reference's SegCVTransformRandomCrop -> RandomFlip -> NormalizeToTensor chain produced for the same seeds
    (tests/golden/input_pipeline.npz, written by the reference's own classes)."""
    import input_recipe as IR
    gold = np.load(os.path.join(HERE, 'golden', 'input_pipeline.npz'))
    case = IR.CASES[name]
    tf, samples, params = _drawn(case)
    got = _statement(case, samples, params, tf.cparams)
    assert np.array_equal(got['image'], gold[name + '.image']) and got['image'].dtype == np.float32
    if name == 'pair_colour':
        kinds = [tuple(op for op, _ in c['ops']) for c in tf.cparams]
        assert any(len(k) == 4 for k in kinds) and any(c['grey'] for c in tf.cparams) and any(len(k) == 0 for k in kinds[1::2])
    if case['labels']:
        assert np.array_equal(got['labels'], gold[name + '.labels']) and got['labels'].dtype == np.int64
    if case['mask']:
        assert np.array_equal(got['mask'], gold[name + '.mask'])
    if name == 'single_padded':
        assert any(p['padded'] for p in params) and any(not p['padded'] for p in params)


def test_crop_entry_table_layout_matches_the_header():
    """72-byte records, field order of b2_crop_entry in include/b200seg.h."""
    from cutmix_semisup_seg_b200.input_pipeline import DeviceCropFlipNormalize
    img = torch.zeros((5, 7, 3), dtype=torch.uint8)
    arr = DeviceCropFlipNormalize.table([dict(image_arr=img)], [dict(pad_top=1, pad_left=2, padded=1, pos=(3, 4), flips=(True, False, True))],
                                        (8, 9))
    assert arr.dtype.itemsize == 72 and arr.dtype.fields['h0'][1] == 24 and arr.dtype.fields['flip_d'][1] == 68
    rec = arr[0]
    assert rec['image'] == img.data_ptr() and rec['labels'] == 0 and (rec['h0'], rec['w0']) == (5, 7)
    assert (rec['pos_y'], rec['pos_x'], rec['crop_h'], rec['crop_w']) == (3, 4, 8, 9) and (rec['flip_x'], rec['flip_y'], rec['flip_d']) == (1, 0, 1)
    with pytest.raises(ValueError, match='square'):
        DeviceCropFlipNormalize((8, 9), hvflip=True)

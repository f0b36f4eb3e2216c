"""CPU: scale / rotation crops of the input pipeline (reference datapipe/seg_transforms_cv.py:169-449).

* The host-side parameter draws of DeviceRandomCropScaleHung / DeviceRandomCropRotateScale + the integer tables + the kernel's
  arithmetic (stated in numpy, tests/geom_recipe.py) reproduce BIT FOR BIT what the reference's own transform classes (cv2.resize /
  cv2.warpAffine -> flip -> normalise) produced for the same seeds (tests/golden/geom_pipeline.npz, oracle/gen_golden.py).
* The restated OpenCV algorithms are also compared with the installed cv2 directly on random shapes / matrices.
The `-m gpu` half (tests/test_zzz_gpu_input.py) runs the CUDA kernel against the same golden bytes."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(__file__)
sys.path.insert(0, HERE)
import geom_recipe as GR  # noqa: E402


@pytest.mark.parametrize('name', sorted(GR.CASES))
def test_geom_algorithm_matches_the_reference_transform_classes(name):
    gold = np.load(os.path.join(HERE, 'golden', 'geom_pipeline.npz'))
    case = GR.CASES[name]
    tf, samples, params = GR.drawn(case)
    got = GR.statement(samples, params, case['crop_size'], case['mean'], case['std'])
    assert got['image'].dtype == np.float32 and np.array_equal(got['image'], gold[name + '.image'])
    if case['labels']:
        assert got['labels'].dtype == np.int64 and np.array_equal(got['labels'], gold[name + '.labels'])
    if case['mask']:
        assert np.array_equal(got['mask'], gold[name + '.mask'])
    xf = np.stack([p['xf_cv'] for p in params])
    assert xf.dtype == gold[name + '.xf_cv'].dtype and np.array_equal(xf, gold[name + '.xf_cv'])
    # the cases exercise what they are meant to
    from cutmix_semisup_seg_b200.input_pipeline import resize_tables, NEAREST, LINEAR
    if name == 'hung_single':
        assert any(p['padded'] for p in params) and any(not p['padded'] for p in params)
        assert any(resize_tables(p['src_size'], case['crop_size'])[1] for p in params)            # the 2x decimation special case
        assert any(p['src_size'][0] < case['crop_size'][0] for p in params)                       # and up-scaling
    if name == 'hung_single_nonuniform':
        assert any(p['src_size'][0] * case['crop_size'][1] != p['src_size'][1] * case['crop_size'][0] for p in params)
    if name == 'rot_single_nolabels':
        assert {p['image_interp'] for p in params} == {NEAREST, LINEAR}
    if name == 'rot_single_labels':
        assert {p['image_interp'] for p in params} == {NEAREST}


def test_restated_opencv_algorithms_match_the_installed_cv2():
    """resize_tables / warp_tables + the integer arithmetic of tests/geom_recipe.py against cv2.resize / cv2.warpAffine themselves:
    nearest and linear, 1 / 3 channels, up- and down-scaling incl. the exact 2x decimation, reflected and constant borders."""
    cv2 = pytest.importorskip('cv2')
    cv2.setNumThreads(0)
    from cutmix_semisup_seg_b200.input_pipeline import NEAREST, LINEAR
    rng = np.random.RandomState(0)
    flag = {NEAREST: cv2.INTER_NEAREST, LINEAR: cv2.INTER_LINEAR}
    sizes = [((int(np.round(33 / f)), int(np.round(41 / f))), (33, 41)) for f in np.arange(5, 16) / 10.0]
    sizes += [((int(rng.randint(8, 200)), int(rng.randint(8, 200))), (int(rng.randint(8, 120)), int(rng.randint(8, 120)))) for _ in range(12)]
    for (sh, sw), (dh, dw) in sizes:
        s = dict(image_arr=rng.randint(0, 256, size=(sh, sw, 3)).astype(np.uint8), mask_arr=rng.randint(0, 256, size=(sh, sw)).astype(np.uint8),
                 labels_arr=rng.randint(0, 21, size=(sh, sw)).astype(np.uint8))
        for ii, mi in ((LINEAR, LINEAR), (LINEAR, NEAREST), (NEAREST, NEAREST)):
            p = dict(mode=0, pad_top=0, pad_left=0, padded=0, pos=(0, 0), src_size=(sh, sw), image_interp=ii, mask_interp=mi)
            rgba, lab, msk = GR.geom_u8(s, p, (dh, dw))
            assert np.array_equal(rgba[..., :3], cv2.resize(s['image_arr'], (dw, dh), interpolation=flag[ii])), ((sh, sw), (dh, dw), ii)
            assert np.array_equal(msk, cv2.resize(s['mask_arr'], (dw, dh), interpolation=flag[mi]))
            assert np.array_equal(lab, cv2.resize(s['labels_arr'], (dw, dh), interpolation=cv2.INTER_NEAREST))
            assert int(rgba[..., 3].min()) == 255
    for trial in range(16):
        sh, sw = (int(v) for v in rng.randint(12, 160, size=2))
        s = dict(image_arr=rng.randint(0, 256, size=(sh, sw, 3)).astype(np.uint8), mask_arr=rng.randint(0, 256, size=(sh, sw)).astype(np.uint8),
                 labels_arr=rng.randint(0, 21, size=(sh, sw)).astype(np.uint8))
        th, sc = rng.uniform(-3.1, 3.1), np.exp(rng.uniform(-0.5, 0.5))
        m = np.array([[np.cos(th) * sc, np.sin(th) * sc, rng.uniform(-40, 40)], [-np.sin(th) * sc, np.cos(th) * sc, rng.uniform(-40, 40)]],
                     dtype=np.float32)
        dh, dw = (int(v) for v in rng.randint(10, 100, size=2))
        for interp in (NEAREST, LINEAR):
            p = dict(mode=1, matrix=m, image_interp=interp, mask_interp=interp)
            rgba, lab, msk = GR.geom_u8(s, p, (dh, dw))
            assert np.array_equal(rgba[..., :3], cv2.warpAffine(s['image_arr'], m, (dw, dh), flags=flag[interp], borderValue=0,
                                                                 borderMode=cv2.BORDER_REFLECT_101)), (trial, interp)
            assert np.array_equal(msk, cv2.warpAffine(s['mask_arr'], m, (dw, dh), flags=flag[interp], borderValue=0,
                                                       borderMode=cv2.BORDER_CONSTANT))
            assert np.array_equal(lab, cv2.warpAffine(s['labels_arr'], m, (dw, dh), flags=cv2.INTER_NEAREST, borderValue=255,
                                                       borderMode=cv2.BORDER_CONSTANT))


def test_geom_entry_table_layout_matches_the_header():
    """88-byte records in the field order of b2_geom_entry (include/b200seg.h); tables: 3 * (h + w) int32 per sample."""
    from cutmix_semisup_seg_b200.input_pipeline import DeviceRandomCropScaleHung, DeviceRandomCropRotateScale, AREA2
    tf = DeviceRandomCropScaleHung((8, 10), rng=np.random.RandomState(0), flip_rng=np.random.RandomState(1))
    dt = tf.entry_dtype()
    assert dt.itemsize == 88 and dt.fields['h0'][1] == 24 and dt.fields['mode'][1] == 32 and dt.fields['tab_off'][1] == 72 and \
        dt.fields['flip_d'][1] == 84
    img = torch.zeros((16, 20, 3), dtype=torch.uint8)
    p = dict(mode=0, pad_top=0, pad_left=0, padded=0, pos=(0, 0), src_size=(16, 20), image_interp=1, mask_interp=0, flips=(True, False, False))
    ent, tab = tf.tables([dict(image_arr=img), dict(image_arr=img)], [p, p])
    assert tab.dtype == np.int32 and tab.shape == (2 * 3 * 18,) and int(ent[1]['tab_off']) == 54
    assert int(ent[0]['image_interp']) == AREA2 and int(ent[0]['mask_interp']) == 0 and int(ent[0]['flip_x']) == 1     # 16x20 -> 8x10
    with pytest.raises(ValueError, match='square'):
        DeviceRandomCropRotateScale((8, 9), hvflip=True)

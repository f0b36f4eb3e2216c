"""mask_gen drop-in vs golden vectors recorded from the unmodified reference (oracle/gen_golden.py)."""
import hashlib
import json
import os

import numpy as np
import pytest

import mask_gen

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'masks.json')))


@pytest.mark.parametrize('case', GOLD, ids=lambda c: 'n{}_{}x{}_seed{}'.format(c['n'], c['shape'][0], c['shape'][1], c['seed']))
def test_generate_params_bit_exact(case):
    kw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in case['kwargs'].items()}
    gen = mask_gen.BoxMaskGenerator(**kw)
    m = gen.generate_params(case['n'], tuple(case['shape']), rng=np.random.RandomState(case['seed']))
    assert m.dtype == np.float64 and m.shape == (case['n'], 1) + tuple(case['shape'])
    assert hashlib.sha256(m.astype(np.float32).tobytes()).hexdigest() == case['sha256']
    assert [float(s) for s in m.reshape(case['n'], -1).sum(axis=1)] == case['sums']
    if 'mask' in case:
        assert np.array_equal(m.reshape(case['n'], -1).astype(int), np.array(case['mask']))


@pytest.mark.parametrize('case', GOLD[:5], ids=lambda c: 'seed{}'.format(c['seed']))
def test_compact_boxes_rasterise_to_the_same_mask(case):
    """generate_boxes (the 4-int form shipped to the GPU) + toggle rasterisation == dense reference masks."""
    kw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in case['kwargs'].items()}
    gen = mask_gen.BoxMaskGenerator(**kw)
    shape = tuple(case['shape'])
    boxes = gen.generate_boxes(case['n'], shape, rng=np.random.RandomState(case['seed']))
    assert boxes.dtype == np.int32 and boxes.shape == (case['n'], gen.n_boxes, 4)
    dense = gen.rasterize_boxes_numpy(boxes, shape, gen.invert)
    assert hashlib.sha256(dense.astype(np.float32).tobytes()).hexdigest() == case['sha256']


def test_survey_known_answer_8x8():
    m = mask_gen.BoxMaskGenerator(0.5, invert=True).generate_params(2, (8, 8), rng=np.random.RandomState(0))
    exp0 = np.zeros((8, 8)); exp0[1:6, 1:7] = 1
    exp1 = np.zeros((8, 8)); exp1[1:6, 2:8] = 1
    assert np.array_equal(m[0, 0], exp0) and np.array_equal(m[1, 0], exp1)


def test_global_rng_default_and_append_to_batch():
    np.random.seed(3)
    a = mask_gen.BoxMaskGenerator((0.25, 0.5)).generate_params(3, (16, 12))
    np.random.seed(3)
    b = mask_gen.BoxMaskGenerator((0.25, 0.5)).generate_params(3, (16, 12), rng=np.random)
    assert np.array_equal(a, b)
    x = np.zeros((3, 3, 16, 12))
    out = mask_gen.BoxMaskGenerator(0.5).append_to_batch(x)
    assert len(out) == 2 and out[1].shape == (3, 1, 16, 12)


def test_add_mask_params_to_batch():
    gen = mask_gen.BoxMaskGenerator(0.5, invert=True)
    batch = [{'image': np.zeros((3, 10, 14), np.float32)} for _ in range(4)]
    out = mask_gen.AddMaskParamsToBatch(gen)(batch)
    assert all(s['mask_params'].dtype == np.float32 and s['mask_params'].shape == (1, 10, 14) for s in out)
    paired = [{'sample0': {'image': np.zeros((3, 6, 7), np.float32)}, 'sample1': {}} for _ in range(2)]
    out = mask_gen.AddMaskParamsToBatch(gen, compact=True)(paired)
    assert all(s['mask_params'].shape == (1, 4) and s['mask_params'].dtype == np.int32 for s in out)


def test_dense_params_pass_through():
    import torch
    gen = mask_gen.BoxMaskGenerator(0.5)
    t = torch.zeros(2, 1, 4, 4)
    assert gen.torch_masks_from_params(t, (4, 4), 'cpu') is t

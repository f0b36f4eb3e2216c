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

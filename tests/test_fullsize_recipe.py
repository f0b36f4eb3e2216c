"""CPU checks of the full-size parity recipe (tests/fullsize_recipe.py): the product's synthetic-weight generator equals the oracle's
bit for bit, the committed golden files exist and carry the recorded scalars, the recipe's inputs are deterministic."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(__file__)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'oracle'))
sys.path.insert(0, HERE)
import torch_oracle as TO  # noqa: E402
import fullsize_recipe as R  # noqa: E402
import mask_gen  # noqa: E402
from cutmix_semisup_seg_b200 import synthetic  # noqa: E402


def test_product_and_oracle_synthetic_weights_agree_bit_for_bit():
    tmpl = {'conv1.weight': torch.empty(8, 3, 7, 7), 'bn1.weight': torch.empty(8), 'bn1.bias': torch.empty(8),
            'bn1.running_mean': torch.empty(8), 'bn1.running_var': torch.empty(8),
            'bn1.num_batches_tracked': torch.zeros((), dtype=torch.long), 'layer1.0.bn3.weight': torch.empty(8),
            'layer1.0.downsample.1.weight': torch.empty(8), 'layer5.conv2d_list.0.weight': torch.empty(5, 8, 3, 3),
            'layer5.conv2d_list.0.bias': torch.empty(5)}
    a = synthetic.synth_state_dict(tmpl, seed=4, logit_gain=7.0, final_keys=['layer5.conv2d_list.0.weight'])
    b = TO.synth_state_dict(tmpl, seed=4, logit_gain=7.0, final_keys=['layer5.conv2d_list.0.weight'])
    assert list(a) == list(b)
    for k in a:
        assert a[k].dtype == b[k].dtype and torch.equal(a[k], b[k]), k


@pytest.mark.parametrize('name', ['cfg3_small', 'cfg2', 'cfg3'])
def test_golden_files_are_committed_and_complete(name):
    z = np.load(os.path.join(HERE, 'golden', 'fullsize_%s.npz' % name))
    cfg = R.CONFIGS[name]
    it = len(np.atleast_1d(z['sup_loss']))
    assert it == cfg['iters']
    for key in ('sup_loss', 'cons_loss', 'conf_rate'):
        v = np.atleast_1d(z[key])
        assert v.shape == (it,) and np.all(np.isfinite(v))
    assert np.all(np.atleast_1d(z['sup_loss']) > 0) and np.all(np.atleast_1d(z['cons_loss']) > 0)
    # the teacher's confidence must straddle the threshold, otherwise the consistency term is vacuous
    assert np.all(np.atleast_1d(z['conf_rate']) > 0.15) and np.all(np.atleast_1d(z['conf_rate']) < 0.85)
    assert z['student_last'].shape[0] == cfg['classes']


def test_recipe_inputs_are_deterministic_and_shaped_like_the_reference_batches():
    cfg = dict(R.CONFIGS['cfg3_small'], n=2, h=32, w=40)
    mg = mask_gen.BoxMaskGenerator(0.5, invert=True)
    (sx, sy), uns = R.batches(cfg, mg, compact_masks=False)
    (sx2, sy2), uns2 = R.batches(cfg, mg, compact_masks=False)
    assert torch.equal(sx, sx2) and torch.equal(sy, sy2) and all(torch.equal(uns[k], uns2[k]) for k in uns)
    assert sx.shape == (2, 3, 32, 40) and sy.shape == (2, 1, 32, 40) and sy.dtype == torch.int64
    assert uns['mask_params'].shape == (2, 1, 32, 40) and uns['um0'].shape == (2, 1, 32, 40)
    assert not torch.equal(uns['ux0_tea'], uns['ux0_stu'])            # paired views (weak / strong)
    dm = R.dropout_masks(cfg)
    assert set(dm) == {'sup', 'tea0', 'tea1', 'stu'} and dm['sup'].shape == (2, 4, 5, 256)
    (_, _), boxes = R.batches(cfg, mg, compact_masks=True)
    dense = TO.box_masks(boxes['mask_params'].numpy(), (32, 40), invert=True)
    assert np.array_equal(dense, uns['mask_params'].numpy())           # compact boxes == the reference's dense masks

"""CPU: the assembled train-time input pipeline (cutmix_semisup_seg_b200.input_pipeline.DeviceTrainPipeline; reference
train_seg_semisup_mask_mt.py:147-179).  The pipeline's parameter draws (geometry and flips from generators SHARED by the supervised
and unsupervised list, colour jitter from torch's generator) combined with the numpy statements of the kernels reproduce bit for bit
what the reference's own transform classes, composed by the script's own lines, produced (tests/golden/train_pipeline.npz).  The
`-m gpu` half runs DeviceTrainPipeline itself on the B200 against the same bytes."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(__file__)
sys.path.insert(0, HERE)
import pipeline_recipe as PR  # noqa: E402


def make_pipeline(case):
    from cutmix_semisup_seg_b200.input_pipeline import DeviceTrainPipeline
    return DeviceTrainPipeline(case['crop_size'], PR.MEAN, PR.STD, rng=np.random.RandomState(case['seed']),
                               flip_rng=np.random.RandomState(case['seed'] + 1), **PR.options(case))


def crops_statement(pipe, samples, params):
    """RGBA uint8 crops + labels / mask planes from the numpy statements of the kernels (geom_recipe / input_recipe)."""
    import geom_recipe as GR
    import input_recipe as IR
    out = []
    for s, p in zip(samples, params):
        if pipe.kind == 'crop':
            r = IR.reference_statement([s], [p], pipe.geom.crop_size, None, None)
            rgb = np.rint(r['image'][0].transpose(1, 2, 0).astype(np.float64) * 255.0).astype(np.uint8)
            probe = IR.reference_statement([dict(image_arr=np.full(s['image_arr'].shape, 255, np.uint8))], [p], pipe.geom.crop_size, None, None)
            alpha = np.rint(probe['image'][0, 0].astype(np.float64) * 255.0).astype(np.uint8)
            out.append((np.concatenate([rgb, alpha[..., None]], axis=2), r.get('labels', [None])[0], r.get('mask', [None])[0]))
        else:
            rgba, lab, msk = GR.geom_u8(s, p, pipe.geom.crop_size)
            f = p['flips']
            out.append((GR.flip(rgba, f), None if lab is None else GR.flip(lab, f)[None].astype(np.int64),
                        None if msk is None else np.multiply(GR.flip(msk, f), 1. / 255, dtype=np.float64)[None].astype(np.float32)))
    return out


def normalise(rgba):
    v = np.multiply(rgba[..., :3], 1. / 255, dtype=np.float64)
    alpha = np.multiply(rgba[..., 3:4], 1. / 255, dtype=np.float64)
    v = (v - np.array(PR.MEAN)[None, None, :] * alpha) / np.array(PR.STD)[None, None, :]
    return v.transpose(2, 0, 1).astype(np.float32)


@pytest.mark.parametrize('name', sorted(PR.CASES))
def test_pipeline_draws_and_kernel_statements_match_the_reference_composition(name):
    import colour_recipe as CR
    gold = np.load(os.path.join(HERE, 'golden', 'train_pipeline.npz'))
    case = PR.CASES[name]
    pipe = make_pipeline(case)
    assert pipe.unsup_paired == PR.options(case)['aug_strong_colour']
    assert pipe.kind == dict(cityscapes='crop', pascal='hung', isic='rot', plain='crop')[name]
    torch.manual_seed(case['torch_seed'])
    for part in ('sup_a', 'unsup', 'sup_b'):
        samples = PR.make_samples(case, part)
        use = [{k: v for k, v in s.items() if k != ('mask_arr' if part != 'unsup' else 'labels_arr')} for s in samples]
        params = [pipe._draw(s) for s in use]
        crops = crops_statement(pipe, use, params)
        key = name + '.' + part
        if part != 'unsup':
            assert np.array_equal(np.stack([normalise(c[0]) for c in crops]), gold[key + '.image'])
            assert np.array_equal(np.stack([c[1] for c in crops]), gold[key + '.labels'])
        elif not pipe.unsup_paired:
            assert np.array_equal(np.stack([normalise(c[0]) for c in crops]), gold[key + '.image'])
            assert np.array_equal(np.stack([c[2] for c in crops]), gold[key + '.mask'])
        else:
            cparams = [pipe.colour.draw() for _ in use]
            assert np.array_equal(np.stack([normalise(c[0]) for c in crops]), gold[key + '.sample0.image'])
            jit = [np.concatenate([CR.apply(c[0][..., :3], cp), c[0][..., 3:]], axis=2) for c, cp in zip(crops, cparams)]
            assert np.array_equal(np.stack([normalise(j) for j in jit]), gold[key + '.sample1.image'])
            for m in ('sample0', 'sample1'):
                assert np.array_equal(np.stack([c[2] for c in crops]), gold[key + '.' + m + '.mask'])
            assert any(cp['ops'] for cp in cparams)


@pytest.mark.gpu
@pytest.mark.parametrize('name', sorted(PR.CASES))
def test_device_train_pipeline_matches_the_reference_composition(name):
    """DeviceTrainPipeline on the B200 (crop / scale / rotation gather, flips, colour jitter, normalise-to-tensor kernels) against the
    reference's composed transform lists on the same seeds."""
    gold = np.load(os.path.join(HERE, 'golden', 'train_pipeline.npz'))
    case = PR.CASES[name]
    dev = torch.device('cuda:0')
    pipe = make_pipeline(case)
    torch.manual_seed(case['torch_seed'])
    for part in ('sup_a', 'unsup', 'sup_b'):
        samples = [{k: torch.from_numpy(v).to(dev) for k, v in s.items()} for s in PR.make_samples(case, part)]
        key = name + '.' + part
        if part != 'unsup':
            out = pipe.sup_batch(samples)
            assert np.array_equal(out['image'].cpu().numpy(), gold[key + '.image'])
            assert out['labels'].dtype == torch.int64 and np.array_equal(out['labels'].cpu().numpy(), gold[key + '.labels'])
        else:
            out = pipe.unsup_batch(samples)
            if pipe.unsup_paired:
                for m in ('sample0', 'sample1'):
                    assert np.array_equal(out[m]['image'].cpu().numpy(), gold[key + '.' + m + '.image']), m
                    assert np.array_equal(out[m]['mask'].cpu().numpy(), gold[key + '.' + m + '.mask'])
            else:
                assert np.array_equal(out['image'].cpu().numpy(), gold[key + '.image'])
                assert np.array_equal(out['mask'].cpu().numpy(), gold[key + '.mask'])


def test_u8_image_source_split_and_samplers():
    """`--dataset synthetic_u8`: decoded-image stand-ins in the reference's sample format, the seeded semi-supervised split and the
    endless random samplers (RepeatSampler(SubsetRandomSampler))."""
    from cutmix_semisup_seg_b200.synthetic import U8ImageSource
    src = U8ImageSource(40, (32, 48), 21, seed=1, device='cpu', n_sup=10, n_unsup=-1, split_seed=12345)
    assert len(src) == 40 and len(src.sup_ndx) == 10 and sorted(src.unsup_ndx.tolist()) == list(range(40))
    assert np.array_equal(src.sup_ndx, np.random.RandomState(12345).permutation(40)[:10])
    sizes = [tuple(s['image_arr'].shape) for s in src.samples]
    assert all(sh[2] == 3 for sh in sizes) and any(sh[0] < 32 or sh[1] < 48 for sh in sizes) and any(sh[0] > 32 and sh[1] > 48 for sh in sizes)
    s0 = src.samples[0]
    assert s0['image_arr'].dtype == torch.uint8 and s0['labels_arr'].shape == s0['image_arr'].shape[:2]
    assert int(s0['labels_arr'][0, 0]) == 255 and int(s0['labels_arr'][-1, -1]) < 21 and int(s0['mask_arr'].min()) == 255
    g = torch.Generator().manual_seed(0)
    it = src.sampler(src.sup_ndx, 4, g)
    seen = [next(it) for _ in range(5)]                       # 20 draws = two passes over the 10 supervised samples
    flat = [i for b in seen for i in b]
    assert all(len(b) == 4 for b in seen) and set(flat) == set(src.sup_ndx.tolist())
    assert sorted(flat[:10]) == sorted(src.sup_ndx.tolist()) and sorted(flat[10:]) == sorted(src.sup_ndx.tolist())
    assert set(src.sup(seen[0])[0]) == {'image_arr', 'labels_arr'} and set(src.unsup(seen[0])[0]) == {'image_arr', 'mask_arr'}
    # the whole pipeline's host side runs on these samples (parameters only; the kernels need the GPU)
    pipe = make_pipeline(PR.CASES['pascal'])
    p = pipe._draw(src.sup(seen[0])[0])
    assert p['mode'] == 0 and len(p['flips']) == 3


def test_synthetic_u8_is_refused_by_entry_points_that_do_not_wire_it():
    import click
    from cutmix_semisup_seg_b200 import train_loop
    with pytest.raises(click.UsageError, match='not wired into this entry point'):
        train_loop.check_dataset('synthetic_u8', u8_supported=False)
    train_loop.check_dataset('synthetic_u8', u8_supported=True)
    with pytest.raises(click.UsageError, match='synthetic_u8'):
        train_loop.check_dataset('pascal_aug', u8_supported=True)


# ------------------------------------------------------------------------------------------ train_seg_semisup_aug_mt.py pairs
def make_aug_pipeline(case):
    from cutmix_semisup_seg_b200.input_pipeline import DeviceTrainPipeline
    return DeviceTrainPipeline(case['crop_size'], PR.MEAN, PR.STD, rng=np.random.RandomState(case['seed']),
                               flip_rng=np.random.RandomState(case['seed'] + 1), script='aug_mt',
                               aug_offset_range=case['aug_offset_range'], aug_free_scale_rot=case['aug_free_scale_rot'], **PR.options(case))


@pytest.mark.parametrize('name', sorted(PR.AUG_CASES))
def test_aug_pair_draws_matrices_and_kernel_statements_match_the_reference_composition(name):
    """script='aug_mt' (train_seg_semisup_aug_mt.py:126-163 + SegCollate._compute_xf_0_to_1): pair parameters, the flips' matrices,
    `xf0_to_1_cv` / `xf0_to_1`, and the crops (numpy statements of the kernels) against the reference's own classes."""
    import colour_recipe as CR
    gold = np.load(os.path.join(HERE, 'golden', 'aug_pipeline.npz'))
    case = PR.AUG_CASES[name]
    pipe = make_aug_pipeline(case)
    torch.manual_seed(case['torch_seed'])
    samples = [{k: v for k, v in s.items() if k != 'labels_arr'} for s in PR.make_samples(case, 'unsup')]
    params, xf01_cv, xf01 = pipe.draw_pairs(samples)
    assert xf01.dtype == np.float32 and np.array_equal(xf01, gold[name + '.xf0_to_1'])
    assert xf01_cv.dtype == gold[name + '.xf0_to_1_cv'].dtype and np.array_equal(xf01_cv, gold[name + '.xf0_to_1_cv'])
    crops = crops_statement(pipe, [s for s in samples for _ in (0, 1)], params)
    c0, c1 = crops[0::2], crops[1::2]
    assert np.array_equal(np.stack([normalise(c[0]) for c in c0]), gold[name + '.sample0.image'])
    if pipe.colour is not None:
        cparams = [pipe.colour.draw() for _ in samples]
        c1 = [(np.concatenate([CR.apply(c[0][..., :3], cp), c[0][..., 3:]], axis=2), c[1], c[2]) for c, cp in zip(c1, cparams)]
    assert np.array_equal(np.stack([normalise(c[0]) for c in c1]), gold[name + '.sample1.image'])
    assert np.array_equal(np.stack([c[2] for c in c0]), gold[name + '.sample0.mask'])
    assert np.array_equal(np.stack([c[2] for c in c1]), gold[name + '.sample1.mask'])
    if name == 'aug_isic':
        assert any(p['flips'][2] for p in params) and any(not p['flips'][2] for p in params)


@pytest.mark.gpu
@pytest.mark.parametrize('name', sorted(PR.AUG_CASES))
def test_device_aug_pair_pipeline_matches_the_reference_composition(name):
    gold = np.load(os.path.join(HERE, 'golden', 'aug_pipeline.npz'))
    case = PR.AUG_CASES[name]
    dev = torch.device('cuda:0')
    pipe = make_aug_pipeline(case)
    torch.manual_seed(case['torch_seed'])
    samples = [{k: torch.from_numpy(v).to(dev) for k, v in s.items()} for s in PR.make_samples(case, 'unsup')]
    out = pipe.unsup_batch(samples)
    for m in ('sample0', 'sample1'):
        assert np.array_equal(out[m]['image'].cpu().numpy(), gold[name + '.' + m + '.image']), m
        assert np.array_equal(out[m]['mask'].cpu().numpy(), gold[name + '.' + m + '.mask']), m
    assert out['xf0_to_1'].dtype == torch.float32 and np.array_equal(out['xf0_to_1'].cpu().numpy(), gold[name + '.xf0_to_1'])

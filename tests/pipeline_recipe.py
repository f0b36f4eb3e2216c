"""Cases of the assembled train-time input pipeline (train_seg_semisup_mask_mt.py:147-179), shared by oracle/gen_golden.py::
gen_train_pipeline (the reference's own transform classes composed like the entry point composes them) and the tests of
cutmix_semisup_seg_b200.input_pipeline.DeviceTrainPipeline."""
import numpy as np

MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
SIZES = [(60, 97), (30, 64), (66, 25), (70, 130), (40, 41), (64, 99), (33, 88), (90, 60)]

# option sets of the reference's recipes (run_cityscapes_experiments.sh:17, run_pascal_aug_experiments.sh, run_isic2017_experiments.sh:18)
CASES = {
    'cityscapes': dict(crop_size=(24, 48), seed=201, torch_seed=11, opts=dict(aug_hflip=True, aug_strong_colour=True)),
    'pascal': dict(crop_size=(33, 33), seed=211, torch_seed=12, opts=dict(aug_hflip=True, aug_scale_hung=True, aug_strong_colour=True)),
    'isic': dict(crop_size=(28, 28), seed=221, torch_seed=13,
                 opts=dict(aug_hflip=True, aug_vflip=True, aug_hvflip=True, aug_max_scale=1.1, aug_rot_mag=45.0, aug_strong_colour=True)),
    'plain': dict(crop_size=(20, 24), seed=231, torch_seed=14, opts=dict()),
}
OPT_DEFAULTS = dict(aug_hflip=False, aug_vflip=False, aug_hvflip=False, aug_scale_hung=False, aug_max_scale=1.0, aug_scale_non_uniform=False,
                    aug_rot_mag=0.0, aug_strong_colour=False, aug_colour_brightness=0.4, aug_colour_contrast=0.4,
                    aug_colour_saturation=0.4, aug_colour_hue=0.1, aug_colour_prob=0.8, aug_colour_greyscale_prob=0.2)


def options(case):
    return dict(OPT_DEFAULTS, **case['opts'])


def make_samples(case, part):
    """Three seeded groups per case: 'sup_a' (image + labels), 'unsup' (image + mask), 'sup_b' -- drawn in this order from the
    pipeline, whose supervised and unsupervised lists share their generators."""
    rs = np.random.RandomState(case['seed'] + {'sup_a': 1000, 'unsup': 2000, 'sup_b': 3000}[part])
    out = []
    for h, w in SIZES:
        s = dict(image_arr=rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8))
        if part == 'unsup':
            m = np.full((h, w), 255, np.uint8)
            m[:, :2] = 0
            s['mask_arr'] = m
        else:
            lab = rs.randint(0, 21, size=(h, w)).astype(np.uint8)
            lab[:1] = 255
            s['labels_arr'] = lab
        out.append(s)
    return out


# train_seg_semisup_aug_mt.py:126-163: every unsupervised sample becomes a pair of differently augmented crops (+ xf0_to_1)
AUG_CASES = {
    'aug_crop': dict(crop_size=(24, 48), seed=301, torch_seed=21, aug_offset_range=8, aug_free_scale_rot=False,
                     opts=dict(aug_hflip=True, aug_strong_colour=True)),
    'aug_hung': dict(crop_size=(32, 32), seed=311, torch_seed=22, aug_offset_range=6, aug_free_scale_rot=False,
                     opts=dict(aug_hflip=True, aug_vflip=True, aug_scale_hung=True)),
    # run_isic2017_experiments.sh:18
    'aug_isic': dict(crop_size=(28, 28), seed=321, torch_seed=23, aug_offset_range=16, aug_free_scale_rot=False,
                     opts=dict(aug_hflip=True, aug_vflip=True, aug_hvflip=True, aug_max_scale=1.1, aug_rot_mag=45.0, aug_strong_colour=True)),
    'aug_free': dict(crop_size=(24, 32), seed=331, torch_seed=24, aug_offset_range=4, aug_free_scale_rot=True,
                     opts=dict(aug_hflip=True, aug_max_scale=1.3, aug_rot_mag=20.0, aug_scale_non_uniform=True)),
}

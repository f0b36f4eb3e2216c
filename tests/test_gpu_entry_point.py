"""-m gpu: the drop-in entry point end to end through its click command (reference train_seg_semisup_mask_mt.py:581-650 ->
:16-530) on synthetic tensors: mean teacher + CutMix, Pi-model + CutOut, the SGD / schedule / ramp-up / ratio variants
(SURVEY.md 8a row V1)."""
import math
import os
import re

import pytest
from click.testing import CliRunner

import train_seg_semisup_ict
import train_seg_semisup_mask_mt as entry

pytestmark = pytest.mark.gpu

BASE = ['--dataset', 'synthetic', '--no_pretrained', '--freeze_bn', '--crop_size', '65,65', '--batch_size', '2',
        '--iters_per_epoch', '2', '--num_epochs', '2', '--learning_rate', '1e-5', '--conf_thresh', '0.5']
CASES = {
    'mean_teacher_cutmix_adam_dl2': ['--arch', 'resnet101_deeplab_imagenet', '--synthetic_classes', '21'],
    'mean_teacher_cutmix_dl3plus_poly': ['--arch', 'resnet101_deeplabv3plus_imagenet', '--synthetic_classes', '19',
                                         '--lr_sched', 'poly', '--cons_loss_fn', 'kld', '--conf_per_pixel'],
    'pi_cutout_sgd_rampup_ratio2': ['--arch', 'resnet101_deeplab_imagenet', '--model', 'pi', '--mask_mode', 'zero',
                                    '--mask_prop_range', '0:1', '--opt_type', 'sgd', '--sgd_nesterov', '--rampup', '3',
                                    '--unsup_batch_ratio', '2', '--cons_loss_fn', 'logits_var', '--aug_strong_colour'],
    'supervised_only': ['--arch', 'resnet101_deeplab_imagenet', '--cons_weight', '0.0', '--lr_sched', 'cosine'],
    # `--dataset synthetic_u8` (the later --dataset wins): uint8 images of assorted sizes through the device input pipeline with
    # the option sets of the reference's recipes (run_pascal_aug_experiments.sh / run_cityscapes_experiments.sh / run_isic2017_...)
    'u8_pascal_recipe_hung_colour': ['--dataset', 'synthetic_u8', '--arch', 'resnet101_deeplab_imagenet', '--aug_hflip',
                                     '--aug_scale_hung', '--aug_strong_colour', '--n_sup', '6'],
    'u8_cityscapes_recipe_dl3plus': ['--dataset', 'synthetic_u8', '--arch', 'resnet101_deeplabv3plus_imagenet', '--synthetic_classes',
                                     '19', '--aug_hflip', '--aug_strong_colour', '--crop_size', '64,96'],
    'u8_isic_recipe_rot_scale_cutout': ['--dataset', 'synthetic_u8', '--arch', 'resnet101_deeplab_imagenet', '--aug_hflip', '--aug_vflip',
                                        '--aug_hvflip', '--aug_max_scale', '1.1', '--aug_rot_mag', '45.0', '--mask_mode', 'zero',
                                        '--mask_prop_range', '0:1'],
}
ICT_CASES = {
    'ict_mean_teacher_dl2': ['--arch', 'resnet101_deeplab_imagenet', '--synthetic_classes', '21', '--ict_alpha', '0.4'],
    'ict_dl3plus_per_pixel_bce': ['--arch', 'resnet101_deeplabv3plus_imagenet', '--synthetic_classes', '19',
                                  '--conf_per_pixel', '--cons_loss_fn', 'bce', '--opt_type', 'sgd', '--rampup', '2'],
    'ict_u8_device_pipeline_colour': ['--dataset', 'synthetic_u8', '--arch', 'resnet101_deeplab_imagenet', '--synthetic_classes', '21',
                                      '--aug_hflip', '--aug_scale_hung', '--aug_strong_colour'],
}


@pytest.mark.parametrize('name', sorted(CASES))
def test_entry_point_runs_on_synthetic_data(name, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)                      # job_helper writes results/<job>/log_<desc>.txt under the cwd
    r = CliRunner().invoke(entry.experiment, BASE + CASES[name] + ['--job_desc', name], catch_exceptions=False)
    assert r.exit_code == 0, r.output
    lines = [l for l in r.output.splitlines() if l.startswith('Epoch ')]
    assert len(lines) == 2, r.output
    for l in lines:
        m = re.search(r'TRAIN clf loss=([-0-9.enainf]+), consistency loss=([-0-9.enainf]+), conf rate=([-0-9.]+)%, VAL mIoU=([-0-9.]+)%', l)
        assert m, l
        sup, cons, conf, miou = (float(x) for x in m.groups())
        assert math.isfinite(sup) and sup > 0.0
        assert math.isfinite(cons) and cons >= 0.0
        assert 0.0 <= conf <= 100.0 and 0.0 <= miou <= 100.0
        if name == 'supervised_only':
            assert cons == 0.0
    assert 'NaN detected' not in r.output
    log = os.path.join('results', 'train_seg_semisup_mask_mt', 'log_{}.txt'.format(name))
    assert os.path.exists(log) and 'Epoch 2' in open(log).read()
    # a finished job is not executed twice (job_helper.py:40-52)
    r2 = CliRunner().invoke(entry.experiment, BASE + CASES[name] + ['--job_desc', name], catch_exceptions=False)
    assert r2.exit_code == 0 and 'already executed' in r2.output


@pytest.mark.parametrize('name', sorted(ICT_CASES))
def test_ict_entry_point_runs_on_synthetic_data(name, tmp_path, monkeypatch):
    """train_seg_semisup_ict.py through its click command."""
    monkeypatch.chdir(tmp_path)
    r = CliRunner().invoke(train_seg_semisup_ict.experiment, BASE + ICT_CASES[name] + ['--job_desc', name],
                           catch_exceptions=False)
    assert r.exit_code == 0, r.output
    lines = [l for l in r.output.splitlines() if l.startswith('Epoch ')]
    assert len(lines) == 2, r.output
    for l in lines:
        m = re.search(r'TRAIN clf loss=([-0-9.enainf]+), consistency loss=([-0-9.enainf]+), conf rate=([-0-9.]+)%, VAL mIoU=([-0-9.]+)%', l)
        assert m, l
        sup, cons, conf, miou = (float(x) for x in m.groups())
        assert math.isfinite(sup) and sup > 0.0 and math.isfinite(cons) and cons >= 0.0
        assert 0.0 <= conf <= 100.0 and 0.0 <= miou <= 100.0
    assert os.path.exists(os.path.join('results', 'train_seg_semisup_ict', 'log_{}.txt'.format(name)))

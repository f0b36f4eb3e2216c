"""-m gpu: augmentation-driven consistency (reference train_seg_semisup_aug_mt.py:275-398) on the B200
kernels: the affine grid-sample kernel against F.affine_grid + F.grid_sample, the fused augmentation-consistency kernel
against the golden vectors produced by the reference's own source lines (tests/golden/aug_block.json) and against the oracle
on the data sets' class counts, full iterations against the oracle's CPU iterations, and the drop-in entry point.

Every test of this file is binding (round 2: the non-strict xfail gates of round 1 are gone)."""
import json
import math
import os
import re
import sys
import warnings

import pytest
import torch
import torch.nn.functional as F

HERE = os.path.dirname(__file__)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'oracle'))
sys.path.insert(0, HERE)
import torch_oracle as TO  # noqa: E402
import ref_step  # noqa: E402
import optim_weight_ema  # noqa: E402
from architectures import network_architectures as na  # noqa: E402
from aug_recipe import aug_inputs, parse_case  # noqa: E402

pytestmark = [pytest.mark.gpu]
dev = torch.device('cuda:0')
GOLD = json.load(open(os.path.join(HERE, 'golden', 'aug_block.json')))


@pytest.fixture(scope='module')
def be():
    from cutmix_semisup_seg_b200 import ops
    return ops.default_backend()


@pytest.mark.parametrize('shape', [(2, 3, 9, 12), (1, 1, 1, 7), (3, 2, 16, 5), (2, 19, 33, 47), (2, 21, 65, 65)])
def test_affine_grid_sample_matches_torch(be, shape):
    n, c, h, w = shape
    g = torch.Generator().manual_seed(h * 100 + w)
    x = torch.randn(shape, generator=g)
    thetas = [torch.tensor([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]]), torch.tensor([[1.0, 0.0, 5.0], [0.0, 1.0, -4.0]]),
              torch.tensor([[0.9, -0.4, 0.1], [0.35, 1.2, -0.2]]), torch.tensor([[-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])]
    for t in thetas:
        theta = t[None].repeat(n, 1, 1)
        theta[0, 0, 2] += 0.03
        for out_hw in (None, (h + 3, max(w - 2, 1))):
            oh, ow = (h, w) if out_hw is None else out_hw
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                want = F.grid_sample(x, F.affine_grid(theta, (n, c, oh, ow), align_corners=True), align_corners=True)
            got = be.affine_grid_sample(x.to(dev), theta.to(dev), out_hw).cpu()
            assert got.shape == want.shape
            assert (got - want).abs().max().item() <= 2e-5 * max(1.0, x.abs().max().item())
    far = torch.tensor([[1.0, 0.0, 5.0], [0.0, 1.0, -4.0]])[None].repeat(n, 1, 1)
    assert be.affine_grid_sample(x.to(dev), far.to(dev)).abs().max().item() == 0.0      # everything outside: zero padding


@pytest.mark.parametrize('key', sorted(k for k, v in GOLD['cases'].items() if 'raises' not in v))
def test_aug_kernel_matches_reference_lines_golden(be, key):
    """loss, confidence rate, |grad| sums recorded from the reference's own loss-block lines; the gradient tensor
    elementwise against the oracle (pinned to the same golden on the CPU)."""
    exp = GOLD['cases'][key]
    fn, tau, pp, rampup = parse_case(key)
    lt, ls0, x0, x1, um0, um1, theta = aug_inputs()
    ramp = 0.25 if rampup > 0 else 1.0
    out4, dls = be.aug_consistency(lt.to(dev), ls0.to(dev), theta.to(dev), um0.to(dev), um1.to(dev), fn, tau, pp, ramp, 1.0)
    grad = (dls * out4[2]).cpu()
    ltol, gtol = (1e-4, 5e-3) if fn == 'bce' else (2e-5, 2e-5)      # bce: fp32 conditioning, see tests/test_gpu_ict.py
    assert float(out4[0]) == pytest.approx(exp['loss'], rel=ltol), key
    if tau > 0:
        assert float(out4[1]) == pytest.approx(exp['conf_rate_acc'], rel=1e-6), key
    assert float(grad.abs().sum()) == pytest.approx(exp['grad_l1'], rel=max(5e-5, gtol)), key
    assert float(grad.abs().max()) == pytest.approx(exp['grad_max'], rel=max(5e-5, gtol)), key
    ls = ls0.clone().requires_grad_(True)
    loss, _ = TO.aug_consistency_loss(lt, ls, theta, um0, um1, fn, tau, pp, ramp_val=0.25, rampup=rampup)
    loss.backward()
    assert (grad - ls.grad).abs().max().item() <= gtol * ls.grad.abs().max().item() + 1e-9, key


@pytest.mark.parametrize('c,shape', [(19, (4, 33, 47)), (21, (2, 65, 65)), (2, (3, 16, 24)), (7, (2, 9, 11))])
@pytest.mark.parametrize('fn', ['var', 'logits_var', 'logits_smoothl1', 'bce', 'kld'])
def test_aug_kernel_matches_oracle_on_class_counts_of_the_data_sets(be, c, shape, fn):
    """All five loss functions of the kernel (logits_var against the sibling scripts' formula: the reference's own branch
    cannot execute, `strict_reference=False`)."""
    from cutmix_semisup_seg_b200 import synthetic
    n, h, w = shape
    g = torch.Generator().manual_seed(100 + c)
    lt = torch.randn((n, c, h, w), generator=g) * 3; ls0 = torch.randn((n, c, h, w), generator=g) * 3
    um0 = torch.rand((n, 1, h, w), generator=g); um0[:, :, :2] = 0
    um1 = torch.rand((n, 1, h, w), generator=g); um1[:, :, :, -2:] = 0
    theta = synthetic.make_aug_batch(n, h, w, c, rot_mag=15.0, max_scale=1.3, offset_range=3.0)['xf0_to_1']
    for tau, pp in ((0.45, False), (0.45, True), (0.0, False)):
        out4, dls = be.aug_consistency(lt.to(dev), ls0.to(dev), theta.to(dev), um0.to(dev), um1.to(dev), fn, tau, pp, 1.0, 0.3)
        ls = ls0.clone().requires_grad_(True)
        loss, conf = TO.aug_consistency_loss(lt, ls, theta, um0, um1, fn, tau, pp, strict_reference=False)
        (loss * 0.3).backward()                                     # train_seg_semisup_aug_mt.py:397-398
        ltol, gtol = (1e-4, 5e-3) if fn == 'bce' else (2e-5, 5e-5)
        if tau > 0:
            assert float(out4[1]) == pytest.approx(float(conf), abs=1.5 / (n * h * w))      # at most one borderline pixel
        assert float(out4[0]) == pytest.approx(float(loss), rel=ltol, abs=1e-9)
        assert float(out4[3]) == pytest.approx(float(loss) * 0.3, rel=ltol, abs=1e-9)
        grad = (dls * out4[2]).cpu()
        assert (grad - ls.grad).abs().max().item() <= gtol * ls.grad.abs().max().item() + 1e-10


def test_aug_kernel_rejects_bad_arguments():
    from cutmix_semisup_seg_b200 import lib as L
    t = torch.zeros((1, 2, 4, 4), device=dev)
    th = torch.zeros((1, 2, 3), device=dev)
    um = torch.ones((1, 1, 4, 4), device=dev)
    part = torch.zeros((3,), device=dev, dtype=torch.float64)
    with pytest.raises(L.B2Error):          # unknown loss function
        L.call('b2_aug_consistency_fwd_bwd', t.data_ptr(), t.data_ptr(), th.data_ptr(), um.data_ptr(), um.data_ptr(),
               t.clone().data_ptr(), part.data_ptr(), 1, 2, 4, 4, 7, 0.5, 0, None)
    with pytest.raises(L.B2Error):          # missing valid mask
        L.call('b2_aug_consistency_fwd_bwd', t.data_ptr(), t.data_ptr(), th.data_ptr(), None, um.data_ptr(),
               t.clone().data_ptr(), part.data_ptr(), 1, 2, 4, 4, 0, 0.5, 0, None)


@pytest.mark.parametrize('batch_trunk,conf_per_pixel', [(True, False), (False, True)])
def test_aug_iterations_match_oracle(batch_trunk, conf_per_pixel):
    """Two full augmentation-consistency iterations (DeepLab v2, frozen BN, Adam with the duplicated group, EMA) vs the
    oracle's CPU iterations: supervised loss 1e-4, consistency loss 5e-3 (3xTF32 logits), post-step weights within Adam's
    +-lr."""
    from cutmix_semisup_seg_b200 import step as step_mod, synthetic
    n, h, w, c, lr = 2, 65, 65, 21, 3e-5
    student = na.seg.get('resnet101_deeplab_imagenet')(c, pretrained=False)
    final = [k for k in student.state_dict() if 'layer5' in k and k.endswith('weight')]
    sd = TO.synth_state_dict(student.state_dict(), seed=3, logit_gain=4.0, final_keys=final)
    student.load_state_dict(sd)
    teacher = na.seg.get('resnet101_deeplab_imagenet')(c, pretrained=False)
    student.to(dev); teacher.to(dev)
    student.b2_precision = teacher.b2_precision = '3xtf32'
    for p in teacher.parameters():
        p.requires_grad = False
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        optim = step_mod.make_optimizer(student, 'adam', lr, fused_kernel=True)
    ema = optim_weight_ema.EMAWeightOptimizer(teacher, student, 0.99)
    student.train(); teacher.train(); student.freeze_batchnorm(); teacher.freeze_batchnorm()
    trainer = step_mod.MeanTeacherStep(student, teacher, optim, ema, None, cons_weight=0.7, conf_thresh=0.5,
                                       conf_per_pixel=conf_per_pixel, batch_trunk=batch_trunk)
    orc = ref_step.OracleMeanTeacher('deeplab2', sd, lr, cons_weight=0.7, conf_thresh=0.5, conf_per_pixel=conf_per_pixel)
    for it in range(2):
        sup = synthetic.make_sup_batch(n, h, w, c, 10 + it)
        uns = synthetic.make_aug_batch(n, h, w, 20 + it)
        out = trainer.step((sup[0].to(dev), sup[1].to(dev)), [{k: v.to(dev) for k, v in uns.items()}])
        s_ref, c_ref, r_ref = orc.step(sup[0], sup[1], dict(uns))
        assert float(out['sup_loss']) == pytest.approx(s_ref, rel=1e-4)
        assert float(out['cons_loss']) == pytest.approx(c_ref, rel=5e-3, abs=1e-7)
        assert float(out['conf_rate']) == pytest.approx(r_ref, abs=2e-3)
    for name, net, ref in (('teacher', teacher, orc.teacher), ('student', student, orc.student)):
        worst = 0.0
        for k, v in net.state_dict().items():
            if v.dtype == torch.float32:
                r = ref[k].detach()
                worst = max(worst, (v.cpu() - r).abs().max().item() / (r.abs().max().item() + 1e-12))
        assert worst < 1.5e-3, (name, worst)


def test_aug_logits_var_fails_like_the_reference():
    from cutmix_semisup_seg_b200 import step as step_mod, synthetic
    n, h, w, c = 1, 33, 33, 21
    student = na.seg.get('resnet101_deeplab_imagenet')(c, pretrained=False).to(dev)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        optim = step_mod.make_optimizer(student, 'adam', 1e-5, fused_kernel=True)
    student.train(); student.freeze_batchnorm()
    trainer = step_mod.MeanTeacherStep(student, student, optim, None, None, cons_loss_fn='logits_var')
    sup = synthetic.make_sup_batch(n, h, w, c, 1, device=dev)
    uns = synthetic.make_aug_batch(n, h, w, 2, device=dev)
    with pytest.raises(NameError):           # train_seg_semisup_aug_mt.py:373
        trainer.step(sup, [uns])


AUG_CASES = {
    'aug_mean_teacher_dl2_rot_scale': ['--arch', 'resnet101_deeplab_imagenet', '--synthetic_classes', '21', '--aug_rot_mag', '10',
                                       '--aug_max_scale', '1.2', '--aug_offset_range', '4'],
    'aug_dl3plus_per_pixel_kld_sgd': ['--arch', 'resnet101_deeplabv3plus_imagenet', '--synthetic_classes', '19',
                                      '--conf_per_pixel', '--cons_loss_fn', 'kld', '--opt_type', 'sgd', '--rampup', '2',
                                      '--aug_offset_range', '6'],
    # `--dataset synthetic_u8`: uint8 images through the device pipeline's pair mode (script='aug_mt'); BASELINE config 4's option set
    # (run_isic2017_experiments.sh:18) on the DenseNet-161 U-Net, and a scale-hung variant on DeepLab v2
    'aug_u8_isic_recipe_denseunet': ['--dataset', 'synthetic_u8', '--arch', 'densenet161unet', '--synthetic_classes', '2', '--crop_size',
                                     '64,64', '--aug_hflip', '--aug_vflip', '--aug_hvflip', '--aug_max_scale', '1.1', '--aug_rot_mag',
                                     '45.0', '--aug_strong_colour'],
    'aug_u8_scale_hung_dl2': ['--dataset', 'synthetic_u8', '--arch', 'resnet101_deeplab_imagenet', '--synthetic_classes', '21',
                              '--aug_hflip', '--aug_scale_hung', '--aug_offset_range', '8'],
}
BASE = ['--dataset', 'synthetic', '--no_pretrained', '--freeze_bn', '--crop_size', '65,65', '--batch_size', '2',
        '--iters_per_epoch', '2', '--num_epochs', '2', '--learning_rate', '1e-5', '--conf_thresh', '0.5']


@pytest.mark.parametrize('name', sorted(AUG_CASES))
def test_aug_entry_point_runs_on_synthetic_data(name, tmp_path, monkeypatch):
    """train_seg_semisup_aug_mt.py through its click command."""
    from click.testing import CliRunner
    import train_seg_semisup_aug_mt
    monkeypatch.chdir(tmp_path)
    r = CliRunner().invoke(train_seg_semisup_aug_mt.experiment, BASE + AUG_CASES[name] + ['--job_desc', name],
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
    assert os.path.exists(os.path.join('results', 'train_seg_semisup_aug_mt', 'log_{}.txt'.format(name)))

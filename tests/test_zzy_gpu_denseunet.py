"""-m gpu: the DenseNet-161 U-Net (`densenet161unet`, reference architectures/denseunet.py; BASELINE config 4 with the
augmentation-consistency loop) on the B200 kernels: the encoder glue of csrc/unet.cu (2x2 average pool forward / backward,
per-channel gradient scaling into a strided prefix) against torch, the network against the fp64 oracle and the golden logits
of the real module (tests/golden/net_denseunet.npz), an augmentation-consistency iteration against the oracle, the entry point.

Every test of this file is binding (round 2: the non-strict xfail gates of round 1 are gone)."""
import math
import os
import re
import sys
import warnings
from collections import OrderedDict

import numpy as np
import pytest
import torch
import torch.nn.functional as F

HERE = os.path.dirname(__file__)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'oracle'))
import torch_oracle as TO  # noqa: E402
import ref_step  # noqa: E402
import optim_weight_ema  # noqa: E402
from architectures import network_architectures as na, denseunet  # noqa: E402

pytestmark = [pytest.mark.gpu]
dev = torch.device('cuda:0')


@pytest.mark.parametrize('shape,ld', [((2, 6, 8, 96), 96), ((1, 5, 7, 48), 384), ((3, 4, 4, 6), 8)])
def test_avgpool_and_channel_scaling_match_torch(shape, ld):
    from cutmix_semisup_seg_b200 import ops
    from cutmix_semisup_seg_b200.acts import Act
    from cutmix_semisup_seg_b200.kernels import ActKernels
    K = ActKernels(ops.default_backend())
    n, h, w, c = shape
    g = torch.Generator().manual_seed(h * w + c)
    x = torch.randn((n, c, h, w), generator=g)
    xa = Act.alloc(n, h, w, c, dev, ld=ld); xa.view4().copy_(x.permute(0, 2, 3, 1).to(dev))
    ya = Act.alloc(n, h // 2, w // 2, c, dev, ld=ld)
    K.avgpool2x2(xa, ya)
    assert torch.allclose(ya.view4().permute(0, 3, 1, 2).cpu(), F.avg_pool2d(x, 2, 2), rtol=1e-6, atol=1e-6)
    dy = torch.randn((n, c, h // 2, w // 2), generator=g)
    da = Act.alloc(n, h // 2, w // 2, c, dev, ld=ld); da.view4().copy_(dy.permute(0, 2, 3, 1).to(dev))
    xg = x.clone().requires_grad_(True)
    F.avg_pool2d(xg, 2, 2).backward(dy)
    dxa = Act.alloc(n, h, w, c, dev, ld=ld)
    K.avgpool2x2_bwd(da, dxa)
    got = dxa.view4().permute(0, 3, 1, 2).cpu().clone()
    assert torch.allclose(got, xg.grad, rtol=1e-6, atol=1e-7)
    K.avgpool2x2_bwd(da, dxa, accumulate=True)
    assert torch.allclose(dxa.view4().permute(0, 3, 1, 2).cpu(), 2 * got, rtol=1e-6, atol=1e-7)
    scale = torch.randn((c,), generator=g)
    K.scale_channels(xa, scale.to(dev), dxa, accumulate=False)
    want = x * scale.view(1, -1, 1, 1)
    assert torch.allclose(dxa.view4().permute(0, 3, 1, 2).cpu(), want, rtol=1e-6, atol=1e-7)
    K.scale_channels(xa, scale.to(dev), dxa, accumulate=True)
    assert torch.allclose(dxa.view4().permute(0, 3, 1, 2).cpu(), 2 * want, rtol=1e-6, atol=1e-6)


def _compare(net, n, h, w, precision, seed=1):
    torch.manual_seed(seed)
    sd = TO.synth_state_dict(net.state_dict(), seed=seed)
    x = torch.randn(n, 3, h, w)
    dm = (torch.rand(n, h, w, 64) > 0.3).float()
    sd64 = OrderedDict((k, v.double().clone() if v.dtype == torch.float32 else v.clone()) for k, v in sd.items())
    for k, p in net.named_parameters():
        if p.requires_grad:
            sd64[k].requires_grad_(True)
    yo = TO.denseunet_forward(sd64, x.double(), backbone_bn_train=False, head_bn_train=True,
                              dropout_masks=[dm.permute(0, 3, 1, 2).double()])
    dy = torch.randn(yo.shape)
    yo.backward(dy.double())
    net.load_state_dict(sd)
    net.to(dev).train()
    net.freeze_batchnorm()
    net.b2_precision = precision
    net.final_dec_drop.inject([dm])
    y = net(x.to(dev))
    assert y.shape == yo.shape and y.dtype == torch.float32 and y.is_contiguous()
    y.backward(dy.to(dev))
    lerr = (y.detach().cpu().double() - yo.detach()).abs().max().item() / yo.abs().max().item()
    errs = []
    for k, p in net.named_parameters():
        g = sd64[k].grad
        if g is None:
            continue
        errs.append((p.grad.detach().cpu().double() - g).abs().max().item() / (g.abs().max().item() + 1e-30))
    # BatchNorm running statistics, RELATIVE to the buffer's magnitude: the decoder of the DenseNet U-Net normalises sums of
    # un-normalised dense-block features whose variance reaches 1e2..1e3, so an absolute bound (as used for the ResNets, whose
    # statistics are O(1)) would measure the data's scale, not the kernels' error
    stat, stat_abs = 0.0, 0.0
    for k, v in net.state_dict().items():
        if 'running' in k:
            ref = sd64[k].detach()
            d = (v.cpu().double() - ref).abs().max().item()
            stat_abs = max(stat_abs, d)
            stat = max(stat, d / (ref.abs().max().item() + 1e-6))
    print('denseunet %s: logits %.2e, grads median %.2e max %.2e, running stats rel %.2e (abs %.2e)' % (
        precision, lerr, sorted(errs)[len(errs) // 2], max(errs), stat, stat_abs))
    return lerr, sorted(errs), stat


def test_shallow_densenet_unet_3xtf32_tight():
    net = denseunet.DenseUNet(denseunet.TVDenseNet(block_config=(2, 2, 3, 2)), 2, mean=None, std=None, pretrained=False)
    lerr, errs, stat = _compare(net, 2, 64, 96, '3xtf32')
    assert lerr < 1e-4
    assert errs[len(errs) // 2] < 3e-2 and errs[-1] < 2.5e-1        # sqrt(forward error) law, tests/test_gpu_nets.py
    assert stat < 1e-4


def test_full_densenet161_unet_3xtf32():
    net = na.seg.get('densenet161unet')(2)
    lerr, errs, stat = _compare(net, 2, 64, 64, '3xtf32')
    assert lerr < 5e-4
    assert len(errs) == 501 and errs[len(errs) // 2] < 1e-1
    assert stat < 1e-3


def test_logits_match_the_reference_module_golden():
    z = np.load(os.path.join(HERE, 'golden', 'net_denseunet.npz'))
    net = na.seg.get('densenet161unet')(2)
    final = [k for k in net.state_dict() if 'final_clf' in k and k.endswith('weight')]
    net.load_state_dict(TO.synth_state_dict(net.state_dict(), seed=1, final_keys=final))
    net.to(dev).train(); net.freeze_batchnorm()
    net.b2_precision = '3xtf32'
    net.final_dec_drop.p = 0.0
    with torch.no_grad():
        y = net(torch.from_numpy(z['x']).to(dev)).cpu().numpy()
    assert np.abs(y - z['logits']).max() <= 5e-4 * np.abs(z['logits']).max()


def test_densenet_unet_aug_consistency_iterations_match_oracle():
    """BASELINE config 4 in small: DenseNet-161 U-Net, 2 classes, augmentation-driven consistency."""
    from cutmix_semisup_seg_b200 import step as step_mod, synthetic
    n, h, w, c, lr = 2, 64, 64, 2, 1e-5
    student = na.seg.get('densenet161unet')(c)
    final = [k for k in student.state_dict() if 'final_clf' in k and k.endswith('weight')]
    sd = TO.synth_state_dict(student.state_dict(), seed=3, logit_gain=4.0, final_keys=final)
    student.load_state_dict(sd)
    teacher = na.seg.get('densenet161unet')(c)
    student.to(dev); teacher.to(dev)
    student.b2_precision = teacher.b2_precision = '3xtf32'
    for p in teacher.parameters():
        p.requires_grad = False
    student.final_dec_drop.p = teacher.final_dec_drop.p = 0.0
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        optim = step_mod.make_optimizer(student, 'adam', lr, fused_kernel=True)
    ema = optim_weight_ema.EMAWeightOptimizer(teacher, student, 0.99)
    student.train(); teacher.train(); student.freeze_batchnorm(); teacher.freeze_batchnorm()
    trainer = step_mod.MeanTeacherStep(student, teacher, optim, ema, None, cons_weight=0.7, conf_thresh=0.5, conf_per_pixel=True)
    orc = ref_step.OracleMeanTeacher('denseunet', sd, lr, cons_weight=0.7, conf_thresh=0.5, conf_per_pixel=True)
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


def test_config4_entry_point_runs_on_synthetic_data(tmp_path, monkeypatch):
    """train_seg_semisup_aug_mt.py --arch densenet161unet (BASELINE config 4 on synthetic ISIC-shaped data: 2 classes)."""
    from click.testing import CliRunner
    import train_seg_semisup_aug_mt as entry
    monkeypatch.chdir(tmp_path)
    args = ['--dataset', 'synthetic', '--freeze_bn', '--crop_size', '64,64', '--batch_size', '2', '--iters_per_epoch', '2',
            '--num_epochs', '2', '--learning_rate', '1e-5', '--conf_thresh', '0.5', '--arch', 'densenet161unet',
            '--synthetic_classes', '2', '--aug_rot_mag', '10', '--aug_max_scale', '1.2', '--aug_offset_range', '4',
            '--job_desc', 'config4']
    r = CliRunner().invoke(entry.experiment, args, catch_exceptions=False)
    assert r.exit_code == 0, r.output
    lines = [l for l in r.output.splitlines() if l.startswith('Epoch ')]
    assert len(lines) == 2, r.output
    for l in lines:
        m = re.search(r'TRAIN clf loss=([-0-9.enainf]+), consistency loss=([-0-9.enainf]+), conf rate=([-0-9.]+)%, VAL mIoU=([-0-9.]+)%', l)
        assert m, l
        sup, cons, conf, miou = (float(x) for x in m.groups())
        assert math.isfinite(sup) and sup > 0.0 and math.isfinite(cons) and cons >= 0.0
        assert 0.0 <= conf <= 100.0 and 0.0 <= miou <= 100.0

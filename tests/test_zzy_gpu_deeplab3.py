"""-m gpu: torchvision's DeepLab v3 in the reference's DeepLabv3Wrapper (`resnet101_deeplabv3_imagenet` / `_coco`, reference
network_architectures.py:75-98) on the B200 kernels: shallow and full networks against the fp64 oracle (3xTF32), the golden
logits of the real module (tests/golden/net_dl3v3.npz), CutMix iterations against the oracle, the entry point.

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

HERE = os.path.dirname(__file__)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'oracle'))
import torch_oracle as TO  # noqa: E402
import ref_step  # noqa: E402
import mask_gen  # noqa: E402
import optim_weight_ema  # noqa: E402
from architectures import network_architectures as na, deeplab3plus  # noqa: E402

pytestmark = [pytest.mark.gpu]
dev = torch.device('cuda:0')


def _compare(net, n, h, w, classes, precision, seed=1):
    torch.manual_seed(seed)
    sd = TO.synth_state_dict(net.state_dict(), seed=seed)
    x = torch.randn(n, 3, h, w)
    dm = (torch.rand(n, -(-h // 8), -(-w // 8), 256) > 0.5).float()
    sd64 = OrderedDict((k, v.double().clone() if v.dtype == torch.float32 else v.clone()) for k, v in sd.items())
    for k, p in net.named_parameters():
        if p.requires_grad:
            sd64[k].requires_grad_(True)
    yo = TO.deeplab3_forward(sd64, x.double(), backbone_bn_train=False, head_bn_train=True,
                             dropout_masks=[dm.permute(0, 3, 1, 2).double()])
    dy = torch.randn(yo.shape)
    yo.backward(dy.double())
    net.load_state_dict(sd)
    net.to(dev).train()
    net.freeze_batchnorm()
    net.b2_precision = precision
    for m in net.modules():
        if type(m).__name__ == 'B2Dropout':
            m.inject([dm])
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
    stat = max([(v.cpu().double() - sd64[k].detach()).abs().max().item() for k, v in net.state_dict().items() if 'running' in k])
    return lerr, sorted(errs), stat


def test_shallow_deeplab3_3xtf32_tight():
    bb = deeplab3plus.ResNetBackbone([1, 1, 2, 1], [False, True, True])
    net = deeplab3plus.DeepLabv3Wrapper(deeplab3plus.DeepLabV3(bb, deeplab3plus.DeepLabHead(2048, 21)))
    lerr, errs, stat = _compare(net, 3, 64, 96, 21, '3xtf32')
    assert lerr < 1e-4
    assert errs[len(errs) // 2] < 3e-2 and errs[-1] < 2.5e-1        # sqrt(forward error) law, tests/test_gpu_nets.py
    assert stat < 1e-4


def test_full_deeplab3_3xtf32():
    net = na.seg.get('resnet101_deeplabv3_imagenet')(21, pretrained=False)
    lerr, errs, stat = _compare(net, 2, 65, 65, 21, '3xtf32')
    assert lerr < 5e-4
    assert len(errs) == 335 and errs[len(errs) // 2] < 1e-1
    assert stat < 1e-3


def test_logits_match_the_wrapped_torchvision_module_golden():
    """tests/golden/net_dl3v3.npz: torchvision's deeplabv3_resnet101 inside the reference's wrapper, train mode with frozen
    backbone BN, dropout off."""
    z = np.load(os.path.join(HERE, 'golden', 'net_dl3v3.npz'))
    net = na.seg.get('resnet101_deeplabv3_imagenet')(21, pretrained=False)
    final = [k for k in net.state_dict() if 'deeplab.classifier.4' in k and k.endswith('weight')]
    net.load_state_dict(TO.synth_state_dict(net.state_dict(), seed=1, final_keys=final))
    net.to(dev).train(); net.freeze_batchnorm()
    net.b2_precision = '3xtf32'
    for m in net.modules():
        if type(m).__name__ == 'B2Dropout':
            m.p = 0.0
    with torch.no_grad():
        y = net(torch.from_numpy(z['x']).to(dev)).cpu().numpy()
    assert np.abs(y - z['logits']).max() <= 5e-4 * np.abs(z['logits']).max()


@pytest.mark.parametrize('batch_trunk', [True, False])
def test_deeplab3_cutmix_iterations_match_oracle(batch_trunk):
    from cutmix_semisup_seg_b200 import step as step_mod, synthetic
    n, h, w, c, lr = 2, 65, 65, 21, 1e-5
    student = na.seg.get('resnet101_deeplabv3_imagenet')(c, pretrained=False)
    final = [k for k in student.state_dict() if 'deeplab.classifier.4' in k and k.endswith('weight')]
    sd = TO.synth_state_dict(student.state_dict(), seed=3, logit_gain=4.0, final_keys=final)
    student.load_state_dict(sd)
    teacher = na.seg.get('resnet101_deeplabv3_imagenet')(c, pretrained=False)
    student.to(dev); teacher.to(dev)
    student.b2_precision = teacher.b2_precision = '3xtf32'
    for p in teacher.parameters():
        p.requires_grad = False
    for net in (student, teacher):
        for m in net.modules():
            if type(m).__name__ == 'B2Dropout':
                m.p = 0.0                          # the dropout draw cannot be shared with the oracle
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        optim = step_mod.make_optimizer(student, 'adam', lr, fused_kernel=True)
    ema = optim_weight_ema.EMAWeightOptimizer(teacher, student, 0.99)
    student.train(); teacher.train(); student.freeze_batchnorm(); teacher.freeze_batchnorm()
    mg = mask_gen.BoxMaskGenerator(0.5, invert=True)
    trainer = step_mod.MeanTeacherStep(student, teacher, optim, ema, mg, cons_weight=0.7, conf_thresh=0.5,
                                       batch_trunk=batch_trunk)
    orc = ref_step.OracleMeanTeacher('deeplab3', sd, lr, cons_weight=0.7, conf_thresh=0.5)
    for it in range(2):
        sup = synthetic.make_sup_batch(n, h, w, c, 10 + it)
        uns = synthetic.make_unsup_batch(n, h, w, 20 + it, mg, compact_masks=True)
        uns_o = dict(uns)
        uns_o['mask_params'] = torch.from_numpy(TO.box_masks(uns['mask_params'].numpy(), (h, w), invert=True))
        out = trainer.step((sup[0].to(dev), sup[1].to(dev)), [{k: v.to(dev) for k, v in uns.items()}])
        s_ref, c_ref, r_ref = orc.step(sup[0], sup[1], uns_o)
        print('dl3 (v3) cutmix iteration %d: sup %.7f vs %.7f (rel %.1e), cons %.6e vs %.6e, conf %.5f vs %.5f' % (
            it, float(out['sup_loss']), s_ref, abs(float(out['sup_loss']) - s_ref) / abs(s_ref), float(out['cons_loss']), c_ref,
            float(out['conf_rate']), r_ref))
        # iteration 0 is pure forward / backward parity (1e-4, the north-star bound); from iteration 1 on the two runs have
        # taken one Adam step each, whose sign-like update of near-zero gradients amplifies rounding differences (~lr per
        # weight): 1e-3
        assert float(out['sup_loss']) == pytest.approx(s_ref, rel=1e-4 if it == 0 else 1e-3)
        assert float(out['cons_loss']) == pytest.approx(c_ref, rel=5e-3, abs=1e-7)
        assert float(out['conf_rate']) == pytest.approx(r_ref, abs=2e-3)
    for name, net, ref in (('teacher', teacher, orc.teacher), ('student', student, orc.student)):
        worst = 0.0
        for k, v in net.state_dict().items():
            if v.dtype == torch.float32:
                r = ref[k].detach()
                worst = max(worst, (v.cpu() - r).abs().max().item() / (r.abs().max().item() + 1e-12))
        assert worst < 1.5e-3, (name, worst)


def test_deeplab3_entry_point_runs_on_synthetic_data(tmp_path, monkeypatch):
    from click.testing import CliRunner
    import train_seg_semisup_mask_mt as entry
    monkeypatch.chdir(tmp_path)
    args = ['--dataset', 'synthetic', '--no_pretrained', '--freeze_bn', '--crop_size', '65,65', '--batch_size', '2',
            '--iters_per_epoch', '2', '--num_epochs', '2', '--learning_rate', '1e-5', '--conf_thresh', '0.5',
            '--arch', 'resnet101_deeplabv3_imagenet', '--synthetic_classes', '21', '--job_desc', 'dl3']
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

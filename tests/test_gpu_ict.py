"""-m gpu: ICT (interpolation consistency training; reference train_seg_semisup_ict.py:306-392) on the
B200 kernels: per-sample mix (bit-exact), the fused ICT consistency kernel against the golden vectors produced by the
reference's own source lines (tests/golden/ict_block.json) and against the oracle on larger inputs, full ICT iterations
against the oracle's CPU iterations."""
import json
import os
import sys
import warnings

import numpy as np
import pytest
import torch

HERE = os.path.dirname(__file__)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'oracle'))
sys.path.insert(0, HERE)
import torch_oracle as TO  # noqa: E402
import ref_step  # noqa: E402
import optim_weight_ema  # noqa: E402
from architectures import network_architectures as na  # noqa: E402
from ict_recipe import ict_inputs, parse_case, factors_of  # noqa: E402

pytestmark = pytest.mark.gpu
dev = torch.device('cuda:0')
GOLD = json.load(open(os.path.join(HERE, 'golden', 'ict_block.json')))


@pytest.fixture(scope='module')
def be():
    from cutmix_semisup_seg_b200 import ops
    return ops.default_backend()


def test_mix_per_sample_is_bit_exact(be):
    g = torch.Generator().manual_seed(5)
    for shape in ((3, 3, 33, 47), (2, 1, 64, 64), (5, 3, 7, 5)):
        a = torch.randn(shape, generator=g); b = torch.randn(shape, generator=g)
        f = torch.tensor(np.random.RandomState(shape[0]).beta(0.4, 0.4, size=(shape[0],)), dtype=torch.float)
        got = be.mix_per_sample(a.to(dev), b.to(dev), f.to(dev)).cpu()
        want = a * (1.0 - f.view(-1, 1, 1, 1)) + b * f.view(-1, 1, 1, 1)        # train_seg_semisup_ict.py:310 (fp32, CPU torch)
        assert torch.equal(got, want)
        want_np = TO.mix(a.numpy(), b.numpy(), np.broadcast_to(f.view(-1, 1, 1, 1).numpy(), shape).astype(np.float32))
        assert np.array_equal(got.numpy(), want_np)


@pytest.mark.parametrize('key', sorted(GOLD['cases']))
def test_ict_kernel_matches_reference_lines_golden(be, key):
    """loss, confidence rate, |grad| sums recorded from the reference's own loss-block lines; the gradient tensor
    elementwise against the oracle (pinned to the same golden on the CPU)."""
    exp = GOLD['cases'][key]
    fn, tau, pp, rampup = parse_case(key)
    l0, l1, ls0, x0, x1, um0, um1 = ict_inputs()
    f = factors_of(exp)
    um = um0 * (1.0 - f) + um1 * f
    ramp = 0.25 if rampup > 0 else 1.0
    out4, dls = be.ict_consistency(l0.to(dev), l1.to(dev), ls0.to(dev), f.reshape(-1).to(dev), um.to(dev), fn, tau, pp, ramp, 1.0)
    grad = (dls * out4[2]).cpu()
    # bce: log(1 - p + 1e-6) and 1 / (p + 1e-6) amplify the last ulp of a saturated soft-max by up to 6 % per term (fp32
    # conditioning of network_architectures.py:115-118; same allowance as the CutMix kernel test): loss within the
    # north-star 1e-4, gradient 5e-3 of its range.  Every other loss function: 2e-5.
    ltol, gtol = (1e-4, 5e-3) if fn == 'bce' else (2e-5, 2e-5)
    assert float(out4[0]) == pytest.approx(exp['loss'], rel=ltol), key
    if tau > 0:
        assert float(out4[1]) == pytest.approx(exp['conf_rate_acc'], rel=1e-6), key
    assert float(grad.abs().sum()) == pytest.approx(exp['grad_l1'], rel=max(5e-5, gtol)), key
    assert float(grad.abs().max()) == pytest.approx(exp['grad_max'], rel=max(5e-5, gtol)), key
    ls = ls0.clone().requires_grad_(True)
    loss, _ = TO.ict_consistency_loss(l0, l1, ls, f, um, fn, tau, pp, ramp_val=0.25, rampup=rampup)
    loss.backward()
    assert (grad - ls.grad).abs().max().item() <= gtol * ls.grad.abs().max().item() + 1e-9, key


@pytest.mark.parametrize('c,shape', [(19, (4, 33, 47)), (21, (2, 65, 65)), (2, (3, 16, 24)), (7, (2, 9, 11))])
@pytest.mark.parametrize('fn', ['var', 'logits_var', 'logits_smoothl1', 'bce', 'kld'])
def test_ict_kernel_matches_oracle_on_class_counts_of_the_data_sets(be, c, shape, fn):
    n, h, w = shape
    g = torch.Generator().manual_seed(100 + c)
    l0 = torch.randn((n, c, h, w), generator=g) * 3; l1 = torch.randn((n, c, h, w), generator=g) * 3
    ls0 = torch.randn((n, c, h, w), generator=g) * 3
    um = torch.rand((n, 1, h, w), generator=g); um[:, :, :2] = 0
    f = torch.tensor(np.random.RandomState(c).beta(0.5, 0.5, size=(n, 1, 1, 1)), dtype=torch.float)
    for tau, pp in ((0.55, False), (0.55, True), (0.0, False)):
        out4, dls = be.ict_consistency(l0.to(dev), l1.to(dev), ls0.to(dev), f.reshape(-1).to(dev), um.to(dev), fn, tau, pp,
                                       1.0, 0.3)
        ls = ls0.clone().requires_grad_(True)
        loss, conf = TO.ict_consistency_loss(l0, l1, ls, f, um, fn, tau, pp)
        (loss * 0.3).backward()                                     # train_seg_semisup_ict.py:389-390
        ltol, gtol = (1e-4, 5e-3) if fn == 'bce' else (2e-5, 5e-5)       # bce: see the golden test above
        assert float(out4[0]) == pytest.approx(float(loss), rel=ltol, abs=1e-9)
        assert float(out4[3]) == pytest.approx(float(loss) * 0.3, rel=ltol, abs=1e-9)
        if tau > 0:
            assert float(out4[1]) == pytest.approx(float(conf), abs=1.5 / (n * h * w))      # at most one borderline pixel
        grad = (dls * out4[2]).cpu()
        assert (grad - ls.grad).abs().max().item() <= gtol * ls.grad.abs().max().item() + 1e-10


def test_ict_per_pixel_mask_needs_the_batch_mean_mask():
    """C ABI error behaviour: the (N,N,1,H,W) broadcast of the reference cannot be computed without confbar."""
    from cutmix_semisup_seg_b200 import lib as L
    t = torch.zeros((1, 2, 4, 4), device=dev)
    f = torch.zeros((1,), device=dev)
    part = torch.zeros((3,), device=dev, dtype=torch.float64)
    with pytest.raises(L.B2Error):
        L.call('b2_ict_consistency_fwd_bwd', t.data_ptr(), t.data_ptr(), t.data_ptr(), f.data_ptr(), None, None,
               t.clone().data_ptr(), part.data_ptr(), 1, 2, 16, 0, 0.5, 1, None)


@pytest.mark.parametrize('batch_trunk,conf_per_pixel', [(True, False), (False, True)])
def test_ict_iterations_match_oracle(batch_trunk, conf_per_pixel):
    """Three full ICT iterations (DeepLab v2, frozen BN, Adam with the duplicated group, EMA) vs the oracle's CPU
    iterations: supervised loss 1e-4, consistency loss 5e-3 (3xTF32 logits), post-step weights within Adam's +-lr."""
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
    trainer = step_mod.MeanTeacherStep(student, teacher, optim, ema, None, cons_weight=0.3, conf_thresh=0.5,
                                       conf_per_pixel=conf_per_pixel, batch_trunk=batch_trunk)
    orc = ref_step.OracleMeanTeacher('deeplab2', sd, lr, cons_weight=0.3, conf_thresh=0.5, conf_per_pixel=conf_per_pixel)
    for it in range(3):
        sup = synthetic.make_sup_batch(n, h, w, c, 10 + it)
        uns = synthetic.make_ict_batch(n, h, w, 20 + it, 0.4)
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

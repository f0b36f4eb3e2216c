"""-m gpu: the loop variants of SURVEY.md 8a row V1 on the CUDA path against the oracle's CPU iterations (round 1 compared them
only on the CPU doubles): `--model pi` (teacher_net is student_net, no EMA; train_seg_semisup_mask_mt.py:110-113),
`--mask_mode zero` (CutOut, :371-401), `--unsup_batch_ratio 2` (:304), `--rampup` without a confidence threshold (:419-420,
:454-455), paired weak / strong views (:313-323), SGD with momentum and weight decay (:95-98), per-pixel confidence masks and
the other consistency-loss functions, `cons_weight == 0` (supervised only, :303)."""
import os
import sys
import warnings

import pytest
import torch

HERE = os.path.dirname(__file__)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'oracle'))
import torch_oracle as TO  # noqa: E402
import ref_step  # noqa: E402
import mask_gen  # noqa: E402
import optim_weight_ema  # noqa: E402
from architectures import network_architectures as na  # noqa: E402

pytestmark = pytest.mark.gpu
dev = torch.device('cuda:0')
KIND, N, H, W, C, LR = 'resnet101_deeplab_imagenet', 2, 65, 65, 21, 3e-5


def _state_gap(net, ref):
    worst = 0.0
    for k, v in net.state_dict().items():
        if v.dtype == torch.float32:
            r = ref[k].detach()
            worst = max(worst, (v.cpu() - r).abs().max().item() / (r.abs().max().item() + 1e-12))
    return worst


def _student(seed):
    student = na.seg.get(KIND)(C, pretrained=False)
    final = [k for k in student.state_dict() if 'layer5' in k and k.endswith('weight')]
    sd = TO.synth_state_dict(student.state_dict(), seed=seed, logit_gain=4.0, final_keys=final)
    student.load_state_dict(sd)
    student.to(dev)
    student.b2_precision = '3xtf32'
    return student, sd


def test_pi_model_cutout_rampup_batch_ratio_sgd_matches_oracle():
    """Pi model + CutOut + ramp-up (no threshold) + two unsupervised batches per labelled batch + paired views + SGD."""
    from cutmix_semisup_seg_b200 import step as step_mod, synthetic
    student, sd = _student(4)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        optim = step_mod.make_optimizer(student, 'sgd', LR * 10, sgd_momentum=0.9, sgd_nesterov=False, sgd_weight_decay=5e-4,
                                        fused_kernel=True)
    student.train(); student.freeze_batchnorm()
    mg = mask_gen.BoxMaskGenerator((0.0, 1.0), invert=True)
    trainer = step_mod.MeanTeacherStep(student, student, optim, None, mg, cons_loss_fn='logits_smoothl1', cons_weight=0.5,
                                       conf_thresh=0.0, rampup=3, mask_mix=False, unsup_batch_ratio=2)
    assert not trainer._can_batch_trunk([None, None])
    orc = ref_step.OracleMeanTeacher('deeplab2', sd, LR * 10, opt_type='sgd', cons_loss_fn='logits_smoothl1', cons_weight=0.5,
                                     conf_thresh=0.0, rampup=3, mask_mix=False, model='pi')
    for it in range(2):
        sup = synthetic.make_sup_batch(N, H, W, C, 70 + it)
        uns = [synthetic.make_unsup_batch(N, H, W, 80 + 2 * it + r, mg, mask_mix=False, paired=True) for r in range(2)]
        uns_o = []
        for u in uns:
            d = dict(u)
            d['mask_params'] = torch.from_numpy(TO.box_masks(u['mask_params'].numpy(), (H, W), invert=True))
            uns_o.append(d)
        ramp = 0.3 + 0.2 * it
        out = trainer.step((sup[0].to(dev), sup[1].to(dev)), [{k: v.to(dev) for k, v in u.items()} for u in uns], ramp_val=ramp)
        s_ref, c_ref, r_ref = orc.step(sup[0], sup[1], uns_o, ramp_val=ramp)
        print('pi / cutout / ramp / ratio 2, iteration %d: sup %.7f vs %.7f, cons %.6e vs %.6e' % (
            it, float(out['sup_loss']), s_ref, float(out['cons_loss']), c_ref))
        assert float(out['sup_loss']) == pytest.approx(s_ref, rel=1e-4)
        assert float(out['cons_loss']) == pytest.approx(c_ref, rel=1e-3, abs=1e-8)          # sum over the two batches (:461)
    # SGD has no sign normalisation: the weights agree closely
    assert _state_gap(student, orc.student) < 2e-4


@pytest.mark.parametrize('loss_fn,per_pixel,mask_mix', [('kld', True, True), ('bce', False, False), ('logits_var', True, False)])
def test_mean_teacher_loss_variants_match_oracle(loss_fn, per_pixel, mask_mix):
    """Mean teacher with the other consistency-loss functions, per-pixel confidence masks, CutMix and CutOut, Adam + EMA."""
    from cutmix_semisup_seg_b200 import step as step_mod, synthetic
    student, sd = _student(5)
    teacher = na.seg.get(KIND)(C, pretrained=False).to(dev)
    teacher.b2_precision = '3xtf32'
    for p in teacher.parameters():
        p.requires_grad = False
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        optim = step_mod.make_optimizer(student, 'adam', LR, fused_kernel=True)
    ema = optim_weight_ema.EMAWeightOptimizer(teacher, student, 0.99)
    student.train(); teacher.train(); student.freeze_batchnorm(); teacher.freeze_batchnorm()
    mg = mask_gen.BoxMaskGenerator(0.5 if mask_mix else (0.0, 1.0), invert=True)
    trainer = step_mod.MeanTeacherStep(student, teacher, optim, ema, mg, cons_loss_fn=loss_fn, cons_weight=0.3, conf_thresh=0.5,
                                       conf_per_pixel=per_pixel, mask_mix=mask_mix)
    orc = ref_step.OracleMeanTeacher('deeplab2', sd, LR, cons_loss_fn=loss_fn, cons_weight=0.3, conf_thresh=0.5,
                                     conf_per_pixel=per_pixel, mask_mix=mask_mix)
    for it in range(2):
        sup = synthetic.make_sup_batch(N, H, W, C, 30 + it)
        uns = synthetic.make_unsup_batch(N, H, W, 40 + it, mg, mask_mix=mask_mix, paired=True)
        uns_o = dict(uns)
        uns_o['mask_params'] = torch.from_numpy(TO.box_masks(uns['mask_params'].numpy(), (H, W), invert=True))
        out = trainer.step((sup[0].to(dev), sup[1].to(dev)), [{k: v.to(dev) for k, v in uns.items()}])
        s_ref, c_ref, r_ref = orc.step(sup[0], sup[1], uns_o)
        print('%s per_pixel=%s mix=%s iteration %d: sup %.7f vs %.7f, cons %.6e vs %.6e, conf %.5f vs %.5f' % (
            loss_fn, per_pixel, mask_mix, it, float(out['sup_loss']), s_ref, float(out['cons_loss']), c_ref, float(out['conf_rate']), r_ref))
        assert float(out['sup_loss']) == pytest.approx(s_ref, rel=1e-4 if it == 0 else 1e-3)
        # 8450 pixels: one pixel crossing the confidence threshold is 1.2e-4 of the rate
        assert float(out['cons_loss']) == pytest.approx(c_ref, rel=5e-3, abs=1e-7)
        assert float(out['conf_rate']) == pytest.approx(r_ref, abs=2e-3)
    assert _state_gap(teacher, orc.teacher) < 1.5e-3 and _state_gap(student, orc.student) < 3e-3


def test_supervised_only_when_cons_weight_is_zero():
    from cutmix_semisup_seg_b200 import step as step_mod, synthetic
    student, sd = _student(6)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        optim = step_mod.make_optimizer(student, 'adam', LR, fused_kernel=True)
    student.train(); student.freeze_batchnorm()
    trainer = step_mod.MeanTeacherStep(student, student, optim, None, None, cons_weight=0.0)
    orc = ref_step.OracleMeanTeacher('deeplab2', sd, LR, cons_weight=0.0, model='pi')
    sup = synthetic.make_sup_batch(N, H, W, C, 90)
    out = trainer.step((sup[0].to(dev), sup[1].to(dev)), [])
    s_ref, c_ref, r_ref = orc.step(sup[0], sup[1], [])
    assert float(out['sup_loss']) == pytest.approx(s_ref, rel=1e-4)
    assert out['cons_loss'] is None and c_ref == 0.0
    assert _state_gap(student, orc.student) < 1.5e-3

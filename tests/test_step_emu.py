"""Host-logic tests of the training iteration without a GPU: `MeanTeacherStep` (step.py) driven with the torch-CPU doubles of
the kernel interfaces (tests/_emu_kernels.py for the networks, tests/_emu_backend.py for the losses / mixes / masks) against
the oracle's iterations (oracle/ref_step.py, pinned to the reference by tests/test_oracle_golden.py).  Covers the CutMix,
CutOut and ICT branches, the batched-trunk and pass-by-pass schedules, and the world-size-2 data-parallel path over gloo.
The numerical parity of the real kernels is the job of the `-m gpu` tests."""
import os
import sys
import warnings

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)

N, H, W, C, LR = 2, 33, 33, 21, 3e-5
KIND = 'resnet101_deeplab_imagenet'


def _install_doubles():
    from _emu_kernels import EmuKernels
    from _emu_backend import EmuBackend
    from cutmix_semisup_seg_b200 import netbase, ops
    k = EmuKernels()
    saved = (netbase.get_kernels, ops.set_default_backend(EmuBackend()))
    netbase.set_kernels_factory(lambda n_split=1: k)
    return saved


def _restore(saved):
    from cutmix_semisup_seg_b200 import netbase, ops
    netbase.set_kernels_factory(saved[0])
    ops.set_default_backend(saved[1])


@pytest.fixture()
def doubles():
    saved = _install_doubles()
    yield
    _restore(saved)


def _build(mode, conf_per_pixel, batch_trunk, dist_group=None, cons_loss_fn='var', **vat):
    import torch_oracle as TO
    import ref_step
    import mask_gen
    from _emu_backend import EmuEMA
    from architectures import network_architectures as na
    from cutmix_semisup_seg_b200 import step as step_mod
    student = na.seg.get(KIND)(C, pretrained=False)
    final = [k for k in student.state_dict() if 'layer5' in k and k.endswith('weight')]
    sd = TO.synth_state_dict(student.state_dict(), seed=3, logit_gain=4.0, final_keys=final)
    student.load_state_dict(sd)
    teacher = na.seg.get(KIND)(C, pretrained=False)
    for p in teacher.parameters():
        p.requires_grad = False
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        optim = step_mod.make_optimizer(student, 'adam', LR)              # torch Adam, per-tensor loop (duplicated group)
    ema = EmuEMA(teacher, student, 0.99)
    student.train(); teacher.train(); student.freeze_batchnorm(); teacher.freeze_batchnorm()
    mg = mask_gen.BoxMaskGenerator(0.5, invert=True)
    trainer = step_mod.MeanTeacherStep(student, teacher, optim, ema, mg, cons_loss_fn=cons_loss_fn, cons_weight=0.7,
                                       conf_thresh=0.5, conf_per_pixel=conf_per_pixel, mask_mix=(mode != 'cut'),
                                       batch_trunk=batch_trunk, dist_group=dist_group, **vat)
    orc = ref_step.OracleMeanTeacher('deeplab2', sd, LR, cons_loss_fn=cons_loss_fn, cons_weight=0.7, conf_thresh=0.5,
                                     conf_per_pixel=conf_per_pixel, mask_mix=(mode != 'cut'), **vat)
    return student, teacher, trainer, orc, mg


def _batches(mode, mg, seed):
    import torch_oracle as TO
    from cutmix_semisup_seg_b200 import synthetic
    sup = synthetic.make_sup_batch(N, H, W, C, 10 + seed)
    if mode == 'ict':
        uns = synthetic.make_ict_batch(N, H, W, 20 + seed, 0.4)
        uns_o = dict(uns)
    elif mode == 'aug':
        uns = synthetic.make_aug_batch(N, H, W, 20 + seed)
        uns_o = dict(uns)
    elif mode == 'vat':
        uns = synthetic.make_vat_batch(N, H, W, 20 + seed, paired=True, with_noise=True)
        uns_o = dict(uns)
    else:
        uns = synthetic.make_unsup_batch(N, H, W, 20 + seed, mg, mask_mix=(mode == 'mix'), compact_masks=True)
        uns_o = dict(uns)
        uns_o['mask_params'] = torch.from_numpy(TO.box_masks(uns['mask_params'].numpy(), (H, W), invert=True))
    return sup, uns, uns_o


def _state_gap(net, ref):
    worst = 0.0
    for k, v in net.state_dict().items():
        if v.dtype == torch.float32:
            r = ref[k].detach()
            worst = max(worst, (v - r).abs().max().item() / (r.abs().max().item() + 1e-12))
    return worst


@pytest.mark.parametrize('mode,conf_per_pixel,batch_trunk,fn', [('mix', False, True, 'var'), ('mix', True, False, 'kld'),
                                                               ('cut', False, True, 'logits_var'), ('ict', True, True, 'var'),
                                                               ('ict', False, False, 'bce'), ('aug', False, True, 'var'),
                                                               ('aug', True, False, 'kld')])
def test_iteration_host_logic_matches_oracle(doubles, mode, conf_per_pixel, batch_trunk, fn):
    student, teacher, trainer, orc, mg = _build(mode, conf_per_pixel, batch_trunk, cons_loss_fn=fn)
    assert trainer._can_batch_trunk([None]) == batch_trunk
    for it in range(2):
        sup, uns, uns_o = _batches(mode, mg, it)
        with torch.no_grad():        # the iteration never uses autograd; the torch-CPU doubles would otherwise record a graph
            out = trainer.step(sup, [uns])
        s_ref, c_ref, r_ref = orc.step(sup[0], sup[1], uns_o)
        assert float(out['sup_loss']) == pytest.approx(s_ref, rel=2e-5)
        assert float(out['cons_loss']) == pytest.approx(c_ref, rel=2e-4, abs=1e-8)
        assert float(out['conf_rate']) == pytest.approx(r_ref, abs=1e-6)
    # Adam normalises gradients: a weight whose tiny gradient changes sign moves by up to +-lr; everything else ~1e-6
    assert _state_gap(student, orc.student) < 1.5e-3
    assert _state_gap(teacher, orc.teacher) < 1.5e-3


@pytest.mark.parametrize('fn,adaptive,from_student,conf_per_pixel', [('kld', False, False, False), ('logits_var', True, True, True)])
def test_vat_iteration_host_logic_matches_oracle(doubles, fn, adaptive, from_student, conf_per_pixel):
    """VAT (train_seg_semisup_vat_mt.py:228-301, 364-452): direction from the input gradient of the direction network in
    eval mode (no parameter gradient), fixed / adaptive radius, then the CutOut-style consistency step; both implementations
    are driven with the same N(0,1) draw."""
    student, teacher, trainer, orc, mg = _build('vat', conf_per_pixel, True, cons_loss_fn=fn, vat_radius=0.5,
                                                adaptive_vat_radius=adaptive, vat_dir_from_student=from_student)
    for it in range(2):
        sup, uns, uns_o = _batches('vat', mg, it)
        assert not trainer._can_batch_trunk([uns])
        with torch.no_grad():
            out = trainer.step(sup, [uns])
        s_ref, c_ref, r_ref = orc.step(sup[0], sup[1], uns_o)
        assert float(out['sup_loss']) == pytest.approx(s_ref, rel=2e-5)
        assert float(out['cons_loss']) == pytest.approx(c_ref, rel=2e-3, abs=1e-8)     # the perturbation amplifies fp32 noise
        assert float(out['conf_rate']) == pytest.approx(r_ref, abs=2.0 / (N * H * W))
    dir_net = student if from_student else teacher
    assert not dir_net.training                    # vat_dir_net.eval() persists (reference :237)
    assert (teacher if from_student else student).training
    assert _state_gap(student, orc.student) < 1.5e-3
    assert _state_gap(teacher, orc.teacher) < 1.5e-3


def test_vat_deeplab3plus_direction_network_stays_in_eval_mode(doubles):
    """DeepLab v3+ has train-mode BatchNorm in its head: the reference leaves the direction network (here the teacher) in
    eval mode after the first VAT direction pass (train_seg_semisup_vat_mt.py:237), so its later forward passes use running
    statistics.  Two iterations against the oracle (which is pinned to the reference's own classes and lines for this case,
    tests/golden/sibling_iterations.json)."""
    import torch_oracle as TO
    import ref_step
    from _emu_backend import EmuEMA
    from architectures import network_architectures as na
    from cutmix_semisup_seg_b200 import step as step_mod, synthetic
    kind, c, lr = 'resnet101_deeplabv3plus_imagenet', 19, 1e-5
    student = na.seg.get(kind)(c, pretrained=False)
    final = [k for k in student.state_dict() if 'classifier.classifier.6' in k and k.endswith('weight')]
    sd = TO.synth_state_dict(student.state_dict(), seed=3, logit_gain=4.0, final_keys=final)
    student.load_state_dict(sd)
    teacher = na.seg.get(kind)(c, pretrained=False)
    for p in teacher.parameters():
        p.requires_grad = False
    for net in (student, teacher):
        for m in net.modules():
            if hasattr(m, 'next_mask'):
                m.p = 0.0                              # the dropout draw cannot be shared with the oracle
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        optim = step_mod.make_optimizer(student, 'adam', lr)
    ema = EmuEMA(teacher, student, 0.99)
    student.train(); teacher.train(); student.freeze_batchnorm(); teacher.freeze_batchnorm()
    trainer = step_mod.MeanTeacherStep(student, teacher, optim, ema, None, cons_loss_fn='kld', cons_weight=0.7, conf_thresh=0.5,
                                       conf_per_pixel=True, vat_radius=0.5, adaptive_vat_radius=True)
    orc = ref_step.OracleMeanTeacher('deeplab3plus', sd, lr, cons_loss_fn='kld', cons_weight=0.7, conf_thresh=0.5,
                                     conf_per_pixel=True, vat_radius=0.5, adaptive_vat_radius=True)
    orc.start_epoch()
    for it in range(2):
        sup = synthetic.make_sup_batch(N, H, W, c, 10 + it)
        uns = synthetic.make_vat_batch(N, H, W, 20 + it, paired=True, with_noise=True)
        with torch.no_grad():
            out = trainer.step(sup, [uns])
        s_ref, c_ref, r_ref = orc.step(sup[0], sup[1], dict(uns))
        assert float(out['sup_loss']) == pytest.approx(s_ref, rel=5e-5)
        assert float(out['cons_loss']) == pytest.approx(c_ref, rel=5e-3, abs=1e-8)
        assert float(out['conf_rate']) == pytest.approx(r_ref, abs=2.0 / (N * H * W))
        assert not teacher.training and student.training
    assert _state_gap(student, orc.student) < 1.5e-3
    assert _state_gap(teacher, orc.teacher) < 1.5e-3


def test_deeplab3_cutmix_iteration_batched_trunk(doubles):
    """CutMix iteration on torchvision's DeepLab v3 (`resnet101_deeplabv3_imagenet`): the batched-trunk schedule hands ONE
    trunk feature (layer4) to the head instead of v3+'s two."""
    import torch_oracle as TO
    import ref_step
    import mask_gen
    from _emu_backend import EmuEMA
    from architectures import network_architectures as na
    from cutmix_semisup_seg_b200 import step as step_mod, synthetic
    kind, c, lr = 'resnet101_deeplabv3_imagenet', 21, 1e-5
    student = na.seg.get(kind)(c, pretrained=False)
    final = [k for k in student.state_dict() if 'deeplab.classifier.4' in k and k.endswith('weight')]
    sd = TO.synth_state_dict(student.state_dict(), seed=3, logit_gain=4.0, final_keys=final)
    student.load_state_dict(sd)
    teacher = na.seg.get(kind)(c, pretrained=False)
    for p in teacher.parameters():
        p.requires_grad = False
    for net in (student, teacher):
        for m in net.modules():
            if hasattr(m, 'next_mask'):
                m.p = 0.0
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        optim = step_mod.make_optimizer(student, 'adam', lr)
    assert [len(g['params']) for g in optim.param_groups] == [0, 335]         # the wrapper carries no `pretraining` tag
    ema = EmuEMA(teacher, student, 0.99)
    student.train(); teacher.train(); student.freeze_batchnorm(); teacher.freeze_batchnorm()
    mg = mask_gen.BoxMaskGenerator(0.5, invert=True)
    trainer = step_mod.MeanTeacherStep(student, teacher, optim, ema, mg, cons_weight=0.7, conf_thresh=0.5)
    assert trainer._can_batch_trunk([None])
    orc = ref_step.OracleMeanTeacher('deeplab3', sd, lr, cons_weight=0.7, conf_thresh=0.5)
    for it in range(2):
        sup = synthetic.make_sup_batch(N, H, W, c, 10 + it)
        uns = synthetic.make_unsup_batch(N, H, W, 20 + it, mg, compact_masks=True)
        uns_o = dict(uns)
        uns_o['mask_params'] = torch.from_numpy(TO.box_masks(uns['mask_params'].numpy(), (H, W), invert=True))
        with torch.no_grad():
            out = trainer.step(sup, [uns])
        s_ref, c_ref, r_ref = orc.step(sup[0], sup[1], uns_o)
        assert float(out['sup_loss']) == pytest.approx(s_ref, rel=5e-5)
        assert float(out['cons_loss']) == pytest.approx(c_ref, rel=2e-3, abs=1e-8)
        assert float(out['conf_rate']) == pytest.approx(r_ref, abs=2.0 / (N * H * W))
    assert _state_gap(student, orc.student) < 1.5e-3
    assert _state_gap(teacher, orc.teacher) < 1.5e-3


def test_resnet50_unet_cutmix_iteration_batched_trunk(doubles):
    """CutMix iteration on the ResNet-50 U-Net (architectures/resunet.py): five trunk features (the skip connections) are
    split per mini-batch by the batched-trunk schedule; the encoder's unused `fc` layer gets no update."""
    import torch_oracle as TO
    import ref_step
    import mask_gen
    from _emu_backend import EmuEMA
    from architectures import network_architectures as na
    from cutmix_semisup_seg_b200 import step as step_mod, synthetic
    kind, c, lr, h, w = 'resnet50unet_imagenet', 11, 1e-5, 32, 64
    student = na.seg.get(kind)(c, pretrained=False)
    final = [k for k in student.state_dict() if 'final_clf' in k and k.endswith('weight')]
    sd = TO.synth_state_dict(student.state_dict(), seed=3, logit_gain=4.0, final_keys=final)
    student.load_state_dict(sd)
    teacher = na.seg.get(kind)(c, pretrained=False)
    for p in teacher.parameters():
        p.requires_grad = False
    for net in (student, teacher):
        for m in net.modules():
            if hasattr(m, 'next_mask'):
                m.p = 0.0                              # the dropout draw cannot be shared with the oracle
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        optim = step_mod.make_optimizer(student, 'adam', lr)
    ema = EmuEMA(teacher, student, 0.99)
    student.train(); teacher.train(); student.freeze_batchnorm(); teacher.freeze_batchnorm()
    mg = mask_gen.BoxMaskGenerator(0.5, invert=True)
    trainer = step_mod.MeanTeacherStep(student, teacher, optim, ema, mg, cons_weight=0.7, conf_thresh=0.5)
    assert trainer._can_batch_trunk([None])
    orc = ref_step.OracleMeanTeacher('resunet', sd, lr, cons_weight=0.7, conf_thresh=0.5)
    fc0 = student.base_model.fc.weight.detach().clone()
    for it in range(2):
        sup = synthetic.make_sup_batch(N, h, w, c, 10 + it)
        uns = synthetic.make_unsup_batch(N, h, w, 20 + it, mg, compact_masks=True)
        uns_o = dict(uns)
        uns_o['mask_params'] = torch.from_numpy(TO.box_masks(uns['mask_params'].numpy(), (h, w), invert=True))
        with torch.no_grad():
            out = trainer.step(sup, [uns])
        s_ref, c_ref, r_ref = orc.step(sup[0], sup[1], uns_o)
        assert float(out['sup_loss']) == pytest.approx(s_ref, rel=5e-5)
        assert float(out['cons_loss']) == pytest.approx(c_ref, rel=2e-3, abs=1e-8)
        assert float(out['conf_rate']) == pytest.approx(r_ref, abs=2.0 / (N * h * w))
    assert torch.equal(student.base_model.fc.weight.detach(), fc0)
    assert _state_gap(student, orc.student) < 1.5e-3
    assert _state_gap(teacher, orc.teacher) < 1.5e-3


def test_densenet161_unet_aug_consistency_iteration(doubles):
    """BASELINE config 4: DenseNet-161 U-Net, 2 classes, augmentation-driven consistency (train_seg_semisup_aug_mt.py) -- one
    iteration of the batched-trunk schedule (five trunk features, four of them concatenation buffers) against the oracle."""
    import torch_oracle as TO
    import ref_step
    from _emu_backend import EmuEMA
    from architectures import network_architectures as na
    from cutmix_semisup_seg_b200 import step as step_mod, synthetic
    kind, c, lr, h, w = 'densenet161unet', 2, 1e-5, 32, 64
    student = na.seg.get(kind)(c)
    final = [k for k in student.state_dict() if 'final_clf' in k and k.endswith('weight')]
    sd = TO.synth_state_dict(student.state_dict(), seed=3, logit_gain=4.0, final_keys=final)
    student.load_state_dict(sd)
    teacher = na.seg.get(kind)(c)
    for p in teacher.parameters():
        p.requires_grad = False
    student.final_dec_drop.p = teacher.final_dec_drop.p = 0.0          # the dropout draw cannot be shared with the oracle
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        optim = step_mod.make_optimizer(student, 'adam', lr)
    ema = EmuEMA(teacher, student, 0.99)
    student.train(); teacher.train(); student.freeze_batchnorm(); teacher.freeze_batchnorm()
    trainer = step_mod.MeanTeacherStep(student, teacher, optim, ema, None, cons_loss_fn='var', cons_weight=0.7, conf_thresh=0.5,
                                       conf_per_pixel=True)
    assert trainer._can_batch_trunk([None])
    orc = ref_step.OracleMeanTeacher('denseunet', sd, lr, cons_loss_fn='var', cons_weight=0.7, conf_thresh=0.5,
                                     conf_per_pixel=True)
    sup = synthetic.make_sup_batch(N, h, w, c, 10)
    uns = synthetic.make_aug_batch(N, h, w, 20)
    with torch.no_grad():
        out = trainer.step(sup, [uns])
    s_ref, c_ref, r_ref = orc.step(sup[0], sup[1], dict(uns))
    assert float(out['sup_loss']) == pytest.approx(s_ref, rel=5e-5)
    assert float(out['cons_loss']) == pytest.approx(c_ref, rel=2e-3, abs=1e-8)
    assert float(out['conf_rate']) == pytest.approx(r_ref, abs=2.0 / (N * h * w))
    assert _state_gap(student, orc.student) < 1.5e-3
    assert _state_gap(teacher, orc.teacher) < 1.5e-3


def test_vat_rejects_loss_functions_the_reference_rejects(doubles):
    student, teacher, trainer, orc, mg = _build('vat', False, True, cons_loss_fn='logits_smoothl1')
    sup, uns, uns_o = _batches('vat', mg, 0)
    with pytest.raises(ValueError, match='Unknown consistency loss function'):       # train_seg_semisup_vat_mt.py:261-262
        with torch.no_grad():
            trainer.step(sup, [uns])
    with pytest.raises(ValueError, match='Unknown consistency loss function'):
        orc.step(sup[0], sup[1], uns_o)


@pytest.mark.parametrize('batch_trunk', [True])
def test_aug_consistency_logits_var_fails_like_the_reference(doubles, batch_trunk):
    """train_seg_semisup_aug_mt.py:373 reads `delta_prob` before assignment: the reference raises on the first unsupervised
    batch (recorded in tests/golden/aug_block.json); so do the oracle and the iteration."""
    student, teacher, trainer, orc, mg = _build('aug', False, batch_trunk, cons_loss_fn='logits_var')
    sup, uns, uns_o = _batches('aug', mg, 0)
    with pytest.raises(NameError):
        with torch.no_grad():
            trainer.step(sup, [uns])
    with pytest.raises(NameError):
        orc.step(sup[0], sup[1], uns_o)


def test_loop_variants_pi_model_cutout_rampup_and_batch_ratio(doubles):
    """`--model pi` (teacher_net is student_net, no EMA; train_seg_semisup_mask_mt.py:110-113) with
    `--mask_mode zero`, `--rampup` and `--unsup_batch_ratio 2` (:304), paired weak / strong views (:313-323)."""
    import torch_oracle as TO
    import ref_step
    import mask_gen
    from architectures import network_architectures as na
    from cutmix_semisup_seg_b200 import step as step_mod, synthetic
    student = na.seg.get(KIND)(C, pretrained=False)
    final = [k for k in student.state_dict() if 'layer5' in k and k.endswith('weight')]
    sd = TO.synth_state_dict(student.state_dict(), seed=4, logit_gain=4.0, final_keys=final)
    student.load_state_dict(sd)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        optim = step_mod.make_optimizer(student, 'sgd', LR * 10, sgd_momentum=0.9, sgd_nesterov=False, sgd_weight_decay=5e-4)
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
        with torch.no_grad():
            out = trainer.step(sup, uns, ramp_val=ramp)
        s_ref, c_ref, r_ref = orc.step(sup[0], sup[1], uns_o, ramp_val=ramp)
        assert float(out['sup_loss']) == pytest.approx(s_ref, rel=2e-5)
        assert float(out['cons_loss']) == pytest.approx(c_ref, rel=2e-4, abs=1e-8)          # sum over the two batches (:461)
    assert _state_gap(student, orc.student) < 1e-4            # SGD: no sign normalisation, weights agree closely


# ---------------------------------------------------------------------------------------------- world size 2 (gloo)
def _dp_worker(rank, world, port, q):
    try:
        os.environ['MASTER_ADDR'] = '127.0.0.1'
        os.environ['MASTER_PORT'] = str(port)
        dist.init_process_group('gloo', rank=rank, world_size=world)
        torch.set_num_threads(2)
        _install_doubles()
        student, teacher, trainer, orc, mg = _build('mix', False, True, dist_group=True)
        sup, uns, uns_o = _batches('mix', mg, 100 + rank)                    # every rank draws its own batches / masks
        with torch.no_grad():
            out = trainer.step(sup, [uns])
        # (1) the gradient every rank applied == mean over ranks of the per-shard oracle gradients (SURVEY.md 8e caveat 1)
        orc.step(sup[0], sup[1], uns_o)                                       # leaves the local gradients in .grad
        names = [k for k, v in orc.student.items() if v.requires_grad and v.grad is not None]
        local = torch.cat([orc.student[k].grad.reshape(-1) for k in names])
        mean = local.clone()
        dist.all_reduce(mean); mean /= world
        params = dict(student.named_parameters())
        got = torch.cat([params[k].grad.reshape(-1) for k in names])         # logical (NCHW) order, like the oracle's
        gerr = (got - mean).abs().max().item() / (mean.abs().max().item() + 1e-30)
        differs = (local - mean).abs().max().item() / (mean.abs().max().item() + 1e-30)
        # (2) parameters stay replicated: identical bits on both ranks after the step
        flat = torch.cat([p.detach().reshape(-1) for p in student.parameters()])
        other = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(other, flat)
        same = bool(torch.equal(other[0], other[1]))
        q.put((rank, gerr, differs, same, float(out['sup_loss'])))
        dist.destroy_process_group()
    except Exception as e:                                                    # surface the failure instead of a queue timeout
        import traceback
        q.put((rank, 'error', traceback.format_exc(), False, repr(e)))


def test_data_parallel_iteration_world2_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for rank, gerr, differs, same, sup in res:
        assert gerr != 'error', differs
        assert gerr < 2e-4                     # averaged gradient == mean of the per-shard reference gradients
        assert differs > 1e-2                  # ... and the shards really had different gradients
        assert same                            # replicated parameters after the optimiser step
    assert res[0][4] != res[1][4]              # different batches per rank (losses are rank-local, never reduced)

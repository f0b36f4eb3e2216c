"""Pins the oracle (oracle/torch_oracle.py, oracle/ref_step.py) to the golden vectors generated from the
unmodified reference by oracle/gen_golden.py."""
import json
import os
import sys
from collections import OrderedDict

import numpy as np
import pytest
import torch

HERE = os.path.dirname(__file__)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'oracle'))
import torch_oracle as TO  # noqa: E402
import ref_step  # noqa: E402
import mask_gen  # noqa: E402
from architectures import network_architectures as na  # noqa: E402

G = os.path.join(HERE, 'golden')


def _synth(kind, classes, seed, gain=1.0):
    net = na.seg.get(kind)(classes) if kind == 'densenet161unet' else na.seg.get(kind)(classes, pretrained=False)
    final = [k for k in net.state_dict() if ('layer5' in k or 'classifier.classifier.6' in k or 'deeplab.classifier.4' in k
                                              or 'final_clf' in k) and k.endswith('weight')]
    return net, TO.synth_state_dict(net.state_dict(), seed=seed, logit_gain=gain, final_keys=final)


def test_loss_block_known_answers():
    gold = json.load(open(os.path.join(G, 'loss_block.json')))
    torch.manual_seed(0)
    N, C, H, W = 2, 5, 6, 6
    l0 = torch.randn(N, C, H, W) * 4; l1 = torch.randn(N, C, H, W) * 4; ls0 = torch.randn(N, C, H, W) * 4
    um0 = torch.ones(N, 1, H, W); um1 = torch.ones(N, 1, H, W); um0[:, :, 0] = 0; um1[:, :, :, 0] = 0.5
    m = torch.tensor(mask_gen.BoxMaskGenerator(0.5, invert=True).generate_params(N, (H, W), rng=np.random.RandomState(0)).astype(np.float32))
    assert [float(s) for s in m.reshape(N, -1).sum(1)] == gold['mask_sums']
    um = um0 * (1 - m) + um1 * m
    for key, exp in gold['cases'].items():
        fn, pp = key.rsplit('_pp', 1)
        ls = ls0.clone().requires_grad_(True)
        loss, conf = TO.consistency_loss(l0, l1, ls, m, um, fn, 0.6, bool(int(pp)))
        loss.backward()
        assert float(loss) == pytest.approx(exp['loss'], rel=1e-6)
        assert float(ls.grad.abs().sum()) == pytest.approx(exp['grad_l1'], rel=1e-6)
        assert float(conf) == pytest.approx(exp['conf_rate'], rel=1e-7)
    torch.manual_seed(1)
    lg = (torch.randn(2, 5, 6, 6) * 2).requires_grad_(True)
    y = torch.randint(0, 5, (2, 1, 6, 6)); y[:, :, 0] = 255
    ce = TO.supervised_loss(lg, y); ce.backward()
    assert float(ce) == pytest.approx(gold['ce']['loss'], rel=1e-6)
    assert float(lg.grad.abs().sum()) == pytest.approx(gold['ce']['grad_l1'], rel=1e-6)
    assert na.sigmoid_rampup(3, 10) == pytest.approx(gold['sigmoid_rampup_3_10'], rel=1e-12)


@pytest.mark.parametrize('tag,kind,classes', [('dl2', 'resnet101_deeplab_imagenet', 21),
                                               ('dl3', 'resnet101_deeplabv3plus_imagenet', 19),
                                               ('dl3v3', 'resnet101_deeplabv3_imagenet', 21),
                                               ('resunet50', 'resnet50unet_imagenet', 11),
                                               ('denseunet', 'densenet161unet', 2)])
def test_functional_nets_match_reference_modules(tag, kind, classes):
    """Forward logits and parameter gradients of the functional oracle == the reference nn.Modules."""
    z = np.load(os.path.join(G, 'net_%s.npz' % tag))
    net, sd = _synth(kind, classes, seed=1)
    for k, p in net.named_parameters():
        if p.requires_grad:
            sd[k].requires_grad_(True)
    x = torch.from_numpy(z['x'])
    if tag == 'dl2':
        y = TO.deeplab2_forward(sd, x)
    elif tag == 'denseunet':
        y = TO.denseunet_forward(sd, x, backbone_bn_train=False, head_bn_train=True, dropout_masks=None)
    elif tag == 'resunet50':
        y = TO.resunet_forward(sd, x, backbone_bn_train=False, head_bn_train=True, dropout_masks=None)
    elif tag == 'dl3v3':       # torchvision's deeplabv3_resnet101 inside the reference's DeepLabv3Wrapper
        y = TO.deeplab3_forward(sd, x, backbone_bn_train=False, head_bn_train=True, dropout_masks=None)
    else:
        y = TO.deeplab3plus_forward(sd, x, backbone_bn_train=False, head_bn_train=True, dropout_masks=None)
    assert np.abs(y.detach().numpy() - z['logits']).max() <= 1e-5 * np.abs(z['logits']).max()
    y.backward(torch.from_numpy(z['dy']))
    names = [str(n) for n in z['grad_names']]
    assert names == [k for k, p in net.named_parameters() if sd[k].grad is not None]
    gabs = np.array([float(sd[k].grad.double().abs().sum()) for k in names])
    assert np.allclose(gabs, z['grad_abs'], rtol=2e-3, atol=1e-6)
    if tag == 'dl3':
        for k in ('deeplab.classifier.project.1.running_mean', 'deeplab.classifier.project.1.running_var'):
            assert np.allclose(sd[k].numpy(), z[k], rtol=1e-5, atol=1e-6)
    if tag in ('resunet50', 'denseunet'):
        for k in ('final_dec_bn.running_mean', 'final_dec_bn.running_var'):
            assert np.allclose(sd[k].numpy(), z[k], rtol=1e-5, atol=1e-6)
    if tag == 'dl3v3':
        for k in ('deeplab.classifier.2.running_mean', 'deeplab.classifier.2.running_var'):
            assert np.allclose(sd[k].numpy(), z[k], rtol=1e-5, atol=1e-6)


def test_iteration_matches_reference_loop():
    """Two iterations of the oracle's loop body == the reference classes driven by the same loop (Adam with the
    duplicated parameter group, EMA incl. BN buffers, CutMix var loss)."""
    gold = json.load(open(os.path.join(G, 'iteration.json')))
    n, h, w, c = gold['n'], gold['h'], gold['w'], gold['classes']
    net, sd = _synth(gold['kind'], c, seed=gold['seed'], gain=gold['gain'])
    tr = ref_step.OracleMeanTeacher('deeplab2', sd, gold['lr'], conf_thresh=gold['conf_thresh'])
    assert [len(g['params']) for g in tr.optim.param_groups] == [314, 8]
    mg = mask_gen.BoxMaskGenerator(prop_range=0.5, n_boxes=1, random_aspect_ratio=True, prop_by_area=True,
                                   within_bounds=True, invert=True)
    for it, exp in enumerate(gold['steps']):
        g = torch.Generator().manual_seed(100 + it)
        sup_x = torch.randn((n, 3, h, w), generator=g)
        sup_y = torch.randint(0, c, (n, 1, h, w), generator=g); sup_y[:, :, :4] = 255
        ux0 = torch.randn((n, 3, h, w), generator=g); ux1 = torch.randn((n, 3, h, w), generator=g)
        um0 = torch.ones((n, 1, h, w)); um0[:, :, :3] = 0; um1 = torch.ones((n, 1, h, w)); um1[:, :, :, 16] = 0.5
        masks = torch.tensor(mg.generate_params(n, (h, w), rng=np.random.RandomState(7 + it)).astype(np.float32))
        uns = dict(ux0_tea=ux0, ux0_stu=ux0, ux1_tea=ux1, ux1_stu=ux1, um0=um0, um1=um1, mask_params=masks)
        s, cl, cr = tr.step(sup_x, sup_y, uns)
        assert s == pytest.approx(exp['sup_loss'], rel=1e-5)
        assert cl == pytest.approx(exp['cons_loss'], rel=2e-3, abs=1e-9)
        assert cr == pytest.approx(exp['conf_rate'], abs=1e-6)
        assert float(tr.student['conv1.weight'].double().sum()) == pytest.approx(exp['student_conv1_sum'], rel=1e-5)
        assert float(tr.student['layer3.5.conv2.weight'].double().sum()) == pytest.approx(exp['student_l3_sum'], rel=1e-5)
        assert float(tr.student['layer5.conv2d_list.0.weight'].double().sum()) == pytest.approx(exp['student_l5_sum'], rel=1e-5)
        assert float(tr.teacher['layer5.conv2d_list.1.weight'].double().sum()) == pytest.approx(exp['teacher_l5_sum'], rel=1e-5)
        tsum = float(sum(v.double().abs().sum() for v in tr.teacher.values() if v.dtype == torch.float32))
        assert tsum == pytest.approx(exp['teacher_abs_sum'], rel=1e-7)


def test_ict_loss_block_matches_the_reference_source_lines():
    """oracle ICT block (torch_oracle.ict_consistency_loss) vs the reference's own lines executed by gen_golden.py,
    including the (N,N,1,H,W) broadcast of the per-pixel confidence mask."""
    from ict_recipe import ict_inputs, parse_case, factors_of
    gold = json.load(open(os.path.join(G, 'ict_block.json')))
    l0, l1, ls0, x0, x1, um0, um1 = ict_inputs()
    assert len(gold['cases']) == 15
    for key, exp in gold['cases'].items():
        fn, tau, pp, rampup = parse_case(key)
        f = factors_of(exp)
        um = um0 * (1.0 - f) + um1 * f
        assert float((x0 * (1.0 - f) + x1 * f).double().sum()) == pytest.approx(exp['mixed_sum'], rel=1e-12)
        assert float(um.double().sum()) == pytest.approx(exp['um_mixed_sum'], rel=1e-12)
        ls = ls0.clone().requires_grad_(True)
        loss, conf = TO.ict_consistency_loss(l0, l1, ls, f, um, fn, tau, pp, ramp_val=0.25, rampup=rampup)
        loss.backward()
        assert float(loss) == pytest.approx(exp['loss'], rel=1e-6), key
        assert float(ls.grad.abs().sum()) == pytest.approx(exp['grad_l1'], rel=1e-6), key
        assert float(ls.grad.abs().max()) == pytest.approx(exp['grad_max'], rel=1e-6), key
        if tau > 0:
            assert float(conf) == pytest.approx(exp['conf_rate_acc'], rel=1e-6), key
        else:
            assert exp['conf_rate_acc'] == 0.25           # `elif rampup > 0: conf_rate_acc += ramp_val` (:352-353)


def test_aug_loss_block_matches_the_reference_source_lines():
    """oracle augmentation-consistency block (torch_oracle.aug_consistency_loss) vs the reference's own lines
    (train_seg_semisup_aug_mt.py:291-394) executed by gen_golden.py -- including the logits_var branch, which the reference
    cannot execute -- and the kernel's algorithm (tests/_emu_backend.py: explicit four-tap gather per student pixel, soft-max
    per tap) against the same vectors."""
    sys.path.insert(0, HERE)
    from aug_recipe import aug_inputs, parse_case
    from _emu_backend import EmuBackend
    import torch.nn.functional as F
    gold = json.load(open(os.path.join(G, 'aug_block.json')))
    lt, ls0, x0, x1, um0, um1, theta = aug_inputs()
    assert len(gold['cases']) == 15
    be = EmuBackend()
    grid = F.affine_grid(theta, x0.shape, align_corners=True)
    for key, exp in gold['cases'].items():
        fn, tau, pp, rampup = parse_case(key)
        if 'raises' in exp:
            assert fn == 'logits_var' and exp['raises'] == 'NameError'
            with pytest.raises(NameError):
                TO.aug_consistency_loss(lt, ls0, theta, um0, um1, fn, tau, pp, ramp_val=0.25, rampup=rampup)
            continue
        assert float((F.grid_sample(um0, grid, align_corners=True) * um1).double().sum()) == pytest.approx(exp['mask_sum'], rel=1e-9)
        ls = ls0.clone().requires_grad_(True)
        loss, conf = TO.aug_consistency_loss(lt, ls, theta, um0, um1, fn, tau, pp, ramp_val=0.25, rampup=rampup)
        loss.backward()
        assert float(loss) == pytest.approx(exp['loss'], rel=1e-6), key
        assert float(ls.grad.abs().sum()) == pytest.approx(exp['grad_l1'], rel=1e-6), key
        assert float(ls.grad.abs().max()) == pytest.approx(exp['grad_max'], rel=1e-6), key
        if tau > 0:
            assert float(conf) == pytest.approx(exp['conf_rate_acc'], rel=1e-6), key
        else:
            assert exp['conf_rate_acc'] == 0.25           # `elif rampup > 0: conf_rate_acc += ramp_val` (:357-358)
        # the kernel's algorithm on the CPU
        ramp = 0.25 if rampup > 0 else 1.0
        out4, dls = be.aug_consistency(lt, ls0, theta, um0, um1, fn, tau, pp, ramp, 1.0)
        grad = dls * out4[2]
        gtol = 5e-3 if fn == 'bce' else 2e-5              # bce: fp32 conditioning of 1 / (p + 1e-6), see tests/test_gpu_ict.py
        assert float(out4[0]) == pytest.approx(exp['loss'], rel=2e-6), key
        if tau > 0:
            assert float(out4[1]) == pytest.approx(exp['conf_rate_acc'], rel=1e-6), key
        assert (grad - ls.grad).abs().max().item() <= gtol * ls.grad.abs().max().item(), key
    # the resampled tensors themselves
    assert float(be.affine_grid_sample(lt, theta).double().abs().sum()) == pytest.approx(exp['logits_in_stu_abs_sum'], rel=1e-6)
    assert float(be.affine_grid_sample(F.softmax(lt, dim=1), theta).double().sum()) == pytest.approx(exp['prob_in_stu_sum'], rel=1e-6)


@pytest.mark.parametrize('shape', [(2, 3, 9, 12), (1, 1, 1, 7), (3, 2, 16, 5), (2, 4, 33, 47)])
def test_affine_grid_sample_algorithm_matches_torch(shape):
    """The sampling arithmetic the CUDA kernels implement (linspace base grid, align_corners un-normalisation, four taps with
    zero padding), stated by tests/_emu_backend.py, against F.affine_grid + F.grid_sample: identity, large shifts (everything
    outside), rotations / anisotropic scales on non-square images, and a different output size."""
    sys.path.insert(0, HERE)
    from _emu_backend import EmuBackend
    import torch.nn.functional as F
    n, c, h, w = shape
    g = torch.Generator().manual_seed(h * 100 + w)
    x = torch.randn(shape, generator=g)
    be = EmuBackend()
    thetas = [torch.tensor([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]]), torch.tensor([[1.0, 0.0, 5.0], [0.0, 1.0, -4.0]]),
              torch.tensor([[0.9, -0.4, 0.1], [0.35, 1.2, -0.2]]), torch.tensor([[-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])]
    for t in thetas:
        theta = t[None].repeat(n, 1, 1)
        theta[0, 0, 2] += 0.03                           # per-sample maps
        for out_hw in (None, (h + 3, max(w - 2, 1))):
            oh, ow = (h, w) if out_hw is None else out_hw
            want = F.grid_sample(x, F.affine_grid(theta, (n, c, oh, ow), align_corners=True), align_corners=True)
            got = be.affine_grid_sample(x, theta, out_hw)
            assert got.shape == want.shape
            assert (got - want).abs().max().item() <= 2e-5 * max(1.0, x.abs().max().item())
    far = torch.tensor([[1.0, 0.0, 5.0], [0.0, 1.0, -4.0]])[None].repeat(n, 1, 1)
    assert be.affine_grid_sample(x, far).abs().max().item() == 0.0


def test_vat_perturbation_matches_the_reference_functions():
    """oracle VAT block (torch_oracle.vat_perturbation) vs the reference's own helper functions
    (train_seg_semisup_vat_mt.py:214-301) executed by gen_golden.py with a small seeded direction network."""
    sys.path.insert(0, HERE)
    from vat_recipe import vat_net, vat_inputs, vat_noise
    gold = json.load(open(os.path.join(G, 'vat_block.json')))
    net = vat_net()
    x, x_hat = vat_inputs()
    assert len(gold['cases']) == 8
    for key, exp in gold['cases'].items():
        fn, adaptive = key.split('_adaptive')
        xp = TO.vat_perturbation(net, x, x_hat, vat_noise(exp['seed'], x.shape), fn, 0.5, bool(int(adaptive)))
        assert not xp.requires_grad and xp.shape == x.shape
        assert float(xp.double().abs().sum()) == pytest.approx(exp['abs_sum'], rel=1e-6), key
        assert [float(v) for v in xp.reshape(3, -1).double().norm(dim=1)] == pytest.approx(exp['norms'], rel=1e-6), key
        assert [float(v) for v in xp[:, :, 3, 5].reshape(-1)] == pytest.approx(exp['probe'], rel=1e-5, abs=1e-7), key
        if not int(adaptive):          # fixed radius: |r_adv| = vat_radius * sqrt(C*H*W) for every sample (:298-301)
            assert exp['norms'] == pytest.approx([0.5 * (3 * 10 * 14) ** 0.5] * 3, rel=1e-6)
    with pytest.raises(ValueError):
        TO.vat_perturbation(net, x, x_hat, vat_noise(1, x.shape), 'logits_smoothl1')


def test_vat_kernel_algorithms_match_torch():
    """The arithmetic of csrc/vat.cu as stated by the test doubles (tests/_emu_backend.py, tests/_emu_kernels.py): per-sample
    norm, adaptive radius, normalise-scale-add and col2im (against the adjoint identity <im2col(x), c> == <x, col2im(c)>)."""
    sys.path.insert(0, HERE)
    from _emu_backend import EmuBackend
    from _emu_kernels import EmuKernels
    from cutmix_semisup_seg_b200.acts import Act
    be, K = EmuBackend(), EmuKernels()
    g = torch.Generator().manual_seed(9)
    x = torch.randn((3, 3, 11, 13), generator=g); e = torch.randn((3, 3, 11, 13), generator=g)
    mag = be.sample_l2norm(e)
    assert torch.allclose(mag, e.reshape(3, -1).norm(dim=1), rtol=1e-6)
    want = x + TO.vat_normalize_eps(e) * 0.37
    assert torch.allclose(be.add_scaled_per_sample(x, e, mag, 0.37), want, rtol=1e-6, atol=1e-7)
    dv = x[:, :, 2:, :] - x[:, :, :-2, :]; dh = x[:, :, :, 2:] - x[:, :, :, :-2]
    rad = 0.5 * torch.sqrt((dv.reshape(3, -1) ** 2).sum(1) + (dh.reshape(3, -1) ** 2).sum(1)) * 0.5
    assert torch.allclose(be.vat_adaptive_radius(x, 0.5), rad, rtol=1e-6)
    assert torch.allclose(be.add_scaled_per_sample(x, e, mag, rad), x + TO.vat_normalize_eps(e) * rad.view(-1, 1, 1, 1), rtol=1e-6, atol=1e-7)
    for kh, stride, pad, dil in ((7, 2, 3, 1), (3, 1, 1, 1), (3, 2, 2, 2)):
        xin = K.nchw_to_act(x, 4)
        oh = (11 + 2 * pad - dil * (kh - 1) - 1) // stride + 1; ow = (13 + 2 * pad - dil * (kh - 1) - 1) // stride + 1
        kpad = (kh * kh * 3 + 31) // 32 * 32
        col = K.im2col(xin, kh, kh, stride, pad, dil, oh, ow, kpad)
        c = col.like(); c.base.copy_(torch.randn(c.base.shape, generator=g))
        dx = xin.like(); dx.base.zero_()
        K.col2im(c, dx, kh, kh, stride, pad, dil, oh, ow, kpad)
        lhs = (col.base.view(-1, kpad)[:, :kh * kh * 3].double() * c.base.view(-1, kpad)[:, :kh * kh * 3].double()).sum()
        rhs = (xin.view4().double() * dx.view4().double()).sum()
        assert float(lhs) == pytest.approx(float(rhs), rel=1e-6)


@pytest.mark.parametrize('run', ['aug', 'vat', 'ict', 'aug_config4'])
def test_sibling_iterations_match_reference_loops(run):
    """Two iterations of the oracle's augmentation-consistency / VAT loop == the reference's classes driven by the reference
    scripts' own unsupervised-branch lines (oracle/gen_golden.py::gen_sibling_iterations): losses, confidence rate, post-step
    student and teacher state, and the train / eval mode the reference leaves the networks in."""
    sys.path.insert(0, HERE)
    from aug_recipe import affine_thetas
    from vat_recipe import vat_noise
    mode = run.split('_')[0]          # 'aug_config4': BASELINE config 4, the aug loop on the DenseNet-161 U-Net
    gold = json.load(open(os.path.join(G, 'sibling_iterations.json')))['runs'][run]
    n, h, w, c = gold['n'], gold['h'], gold['w'], gold['classes']
    net, sd = _synth(gold['kind'], c, seed=gold['seed'], gain=gold['gain'])
    arch = {'vat': 'deeplab3plus', 'aug_config4': 'denseunet'}.get(run, 'deeplab2')
    tr = ref_step.OracleMeanTeacher(arch, sd, gold['lr'], cons_loss_fn=gold['cons_loss_fn'], cons_weight=gold['cons_weight'],
                                    conf_thresh=gold['conf_thresh'], conf_per_pixel=gold['conf_per_pixel'],
                                    vat_radius=gold['vat_radius'], adaptive_vat_radius=gold['adaptive_vat_radius'])
    tr.start_epoch()
    conv1 = {'vat': 'deeplab.backbone.conv1.weight', 'aug_config4': 'base_model.features.conv0.weight'}.get(run, 'conv1.weight')
    for it, exp in enumerate(gold['steps']):
        g = torch.Generator().manual_seed(300 + it)
        sup_x = torch.randn((n, 3, h, w), generator=g)
        sup_y = torch.randint(0, c, (n, 1, h, w), generator=g); sup_y[:, :, :4] = 255
        ux0 = torch.randn((n, 3, h, w), generator=g); ux1 = ux0 + 0.1 * torch.randn((n, 3, h, w), generator=g)
        um0 = torch.ones((n, 1, h, w)); um0[:, :, :3] = 0; um1 = torch.ones((n, 1, h, w)); um1[:, :, :, 16] = 0.5
        if mode == 'aug':
            theta = affine_thetas()[it:it + 2] if it == 0 else affine_thetas()[[2, 0]]
            uns = dict(ux0=ux0, ux1=ux1, um0=um0, um1=um1, xf0_to_1=theta)
        elif mode == 'ict':
            np.random.seed(700 + it)
            f = torch.tensor(np.random.beta(0.4, 0.4, size=(n, 1, 1, 1)), dtype=torch.float)       # ict :306-307
            assert [float(v) for v in f.reshape(-1)] == exp['factors']
            uns = dict(ux0_tea=ux0, ux0_stu=ux0, ux1_tea=ux1, ux1_stu=ux1, um0=um0, um1=um1, ict_mix_factors=f)
        else:
            uns = dict(ux_tea=ux0, ux_stu=ux1, um=um0, vat=torch.ones(1), noise=vat_noise(500 + it, ux0.shape))
        s, cl, cr = tr.step(sup_x, sup_y, uns)
        assert s == pytest.approx(exp['sup_loss'], rel=1e-5)
        assert cl == pytest.approx(exp['cons_loss'], rel=2e-3, abs=1e-9)
        assert cr == pytest.approx(exp['conf_rate'], abs=1e-6)
        assert tr.eval_mode == {'student': not exp['student_training'], 'teacher': not exp['teacher_training']}
        assert float(tr.student[conv1].double().sum()) == pytest.approx(exp['student_conv1_sum'], rel=1e-5)
        ssum = float(sum(v.double().abs().sum() for v in tr.student.values() if v.dtype == torch.float32))
        tsum = float(sum(v.double().abs().sum() for v in tr.teacher.values() if v.dtype == torch.float32))
        assert ssum == pytest.approx(exp['student_abs_sum'], rel=1e-7)
        assert tsum == pytest.approx(exp['teacher_abs_sum'], rel=1e-7)


def test_bit_exact_elementwise_oracles():
    rs = np.random.RandomState(0)
    t = rs.randn(100003).astype(np.float32); s = rs.randn(100003).astype(np.float32)
    tt = torch.from_numpy(t.copy()); tt.mul_(0.99); tt.add_(torch.from_numpy(s) * (1.0 - 0.99))
    assert np.array_equal(TO.ema_update(t, s, 0.99), tt.numpy())
    a = rs.randn(2, 3, 9, 7).astype(np.float32); b = rs.randn(2, 3, 9, 7).astype(np.float32)
    m = (rs.rand(2, 1, 9, 7) > 0.5).astype(np.float32); m[0, 0, 0, 0] = 0.25
    ref = torch.from_numpy(a) * (1 - torch.from_numpy(m)) + torch.from_numpy(b) * torch.from_numpy(m)
    assert np.array_equal(TO.mix(a, b, m), ref.numpy())

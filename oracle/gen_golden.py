"""Generates tests/golden/*.json|npz by running the UNMODIFIED reference (imported from /root/reference, which
only exists in the development container) — the committed fixtures pin oracle/ and the product on the GPU box.

    cd /tmp && python /root/repo/oracle/gen_golden.py            # writes into /root/repo/tests/golden

Contents
  masks.json        BoxMaskGenerator.generate_params known answers (sha256 of the float32 bytes, means, sums)
  state_dicts.json  state_dict keys / shapes / dtypes of the two hot-path architectures + optimiser group sizes
  net_dl2.npz, net_dl3.npz   logits and per-parameter gradient checksums of the reference modules on seeded
                    synthetic weights / inputs (small crops)
  iteration.json    two full iterations of the loop body (train_seg_semisup_mask_mt.py:287-476) driven with the
                    reference modules, reference EMAWeightOptimizer, torch Adam on the reference param groups
  loss_block.json   known answers for the consistency / CE block recorded in SURVEY.md §8c
  ict_block.json    ICT loss block (train_seg_semisup_ict.py:306-387): the reference's own source lines executed on seeded
                    tensors, all five loss functions x {scalar / per-pixel confidence mask, ramp-up without threshold}
  aug_block.json    augmentation-consistency loss block (train_seg_semisup_aug_mt.py:291-394), the same way
  vat_block.json    VAT perturbation (train_seg_semisup_vat_mt.py:214-301): the reference's own helper functions executed
                    with a small seeded direction network, four loss functions x fixed / adaptive radius
  sibling_iterations.json  two full iterations of the aug (DeepLab v2) and VAT (DeepLab v3+) loops: reference classes + the
                    reference scripts' own unsupervised-branch lines, Adam, EMA (pins oracle/ref_step.py's aug / VAT branches)
  entry_point.json  click surface of the reference's `train_seg_semisup_mask_mt.experiment` (option names, flags,
                    defaults, choices) and the parameter list of the job function, plus lr_schedules / sigmoid_rampup
                    known answers the entry point depends on
"""
import hashlib
import json
import math
import os
import sys
import warnings

REF = os.environ.get('CUTMIX_REF', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
sys.path.insert(0, REF)            # reference packages: architectures, mask_gen, optim_weight_ema
sys.path.insert(1, HERE)           # torch_oracle (synthetic weights, shared with the tests)
warnings.filterwarnings('ignore')

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import mask_gen  # noqa: E402  (reference)
import optim_weight_ema  # noqa: E402  (reference)
from architectures import network_architectures  # noqa: E402  (reference)
import torch_oracle as TO  # noqa: E402

assert os.path.realpath(mask_gen.__file__).startswith(os.path.realpath(REF)), 'reference not first on sys.path'

MASK_CASES = [
    (dict(prop_range=0.5, invert=True), 16, (512, 512), 12345),
    (dict(prop_range=0.5, invert=True), 10, (321, 321), 12345),
    (dict(prop_range=(0.0, 1.0), invert=True), 10, (321, 321), 12345),
    (dict(prop_range=(0.25, 0.5), n_boxes=3, random_aspect_ratio=False, prop_by_area=False, within_bounds=False,
          invert=False), 4, (64, 48), 7),
    (dict(prop_range=0.5, invert=True), 2, (8, 8), 0),
    (dict(prop_range=(0.1, 0.9), n_boxes=2, random_aspect_ratio=False, prop_by_area=True, within_bounds=False,
          invert=True), 6, (33, 47), 3),
    (dict(prop_range=(0.0, 0.3), n_boxes=4, random_aspect_ratio=True, prop_by_area=False, within_bounds=True,
          invert=False), 5, (20, 31), 5),
    (dict(prop_range=(0.5, 1.0), n_boxes=2, random_aspect_ratio=True, prop_by_area=True, within_bounds=False,
          invert=False), 7, (17, 9), 11),
]


def gen_masks():
    out = []
    for kw, n, shape, seed in MASK_CASES:
        m = mask_gen.BoxMaskGenerator(**kw).generate_params(n, shape, rng=np.random.RandomState(seed))
        m32 = m.astype(np.float32)
        rec = dict(kwargs={k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()}, n=n, shape=list(shape), seed=seed,
                   dtype=str(m.dtype), mean=float(m.mean()), sums=[float(s) for s in m.reshape(n, -1).sum(axis=1)],
                   sha256=hashlib.sha256(m32.tobytes()).hexdigest())
        if m.size <= 4096:
            rec['mask'] = m32.reshape(n, -1).astype(int).tolist()
        out.append(rec)
    json.dump(out, open(os.path.join(OUT, 'masks.json'), 'w'), indent=1)


def build(kind, classes, seed, gain=1.0):
    if kind == 'resnet101_deeplabv3_imagenet':
        net = reference_deeplabv3(classes)
    elif kind == 'densenet161unet':
        net = network_architectures.seg.get(kind)(classes)
    else:
        net = network_architectures.seg.get(kind)(classes, pretrained=False)
    final = [k for k in net.state_dict() if ('layer5' in k or 'classifier.classifier.6' in k or 'deeplab.classifier.4' in k
                                              or 'final_clf' in k) and k.endswith('weight')]
    sd = TO.synth_state_dict(net.state_dict(), seed=seed, logit_gain=gain, final_keys=final)
    net.load_state_dict(sd)
    return net


def reference_deeplabv3(classes):
    """The reference's `resnet101_deeplabv3_imagenet(num_classes, pretrained=False)` (network_architectures.py:87-98) =
    torchvision's deeplabv3_resnet101 inside the reference's DeepLabv3Wrapper.  The factory itself cannot run offline: every
    torchvision release downloads ImageNet backbone weights unless told otherwise, so the two constructor calls are made
    here with the backbone download switched off."""
    from torchvision.models.segmentation import deeplabv3_resnet101
    from architectures import deeplab3plus as ref_dl3
    deeplab = deeplabv3_resnet101(weights=None, weights_backbone=None, num_classes=classes)
    return ref_dl3.DeepLabv3Wrapper(deeplab)


def gen_state_dicts():
    out = {}
    for kind, classes in (('resnet101_deeplab_imagenet', 21), ('resnet101_deeplabv3plus_imagenet', 19),
                          ('resnet101_deeplabv3_imagenet', 21), ('resnet50unet_imagenet', 11),
                          ('resnet101unet_imagenet', 11), ('densenet161unet', 2)):
        if kind == 'resnet101_deeplabv3_imagenet':
            net = reference_deeplabv3(classes)
        elif kind == 'densenet161unet':
            net = network_architectures.seg.get(kind)(classes)           # the factory takes no `pretrained` argument
        else:
            net = network_architectures.seg.get(kind)(classes, pretrained=False)
        out[kind] = dict(classes=classes,
                         entries=[[k, list(v.shape), str(v.dtype).replace('torch.', '')] for k, v in net.state_dict().items()],
                         n_pretrained=len(list(net.pretrained_parameters())),
                         n_pretrained_unique=len(set(id(p) for p in net.pretrained_parameters())),
                         n_new=len(list(net.new_parameters())),
                         trainable=[k for k, p in net.named_parameters() if p.requires_grad],
                         registry_names=sorted(network_architectures.seg.names()))
    json.dump(out, open(os.path.join(OUT, 'state_dicts.json'), 'w'))


def gen_nets():
    for tag, kind, classes, (n, h, w) in (('dl2', 'resnet101_deeplab_imagenet', 21, (2, 33, 41)),
                                         ('dl3', 'resnet101_deeplabv3plus_imagenet', 19, (3, 33, 41)),
                                         ('dl3v3', 'resnet101_deeplabv3_imagenet', 21, (2, 33, 41)),
                                         ('resunet50', 'resnet50unet_imagenet', 11, (2, 32, 64)),
                                         ('denseunet', 'densenet161unet', 2, (2, 32, 64))):
        net = build(kind, classes, seed=1)
        net.train()
        net.freeze_batchnorm()
        for m in net.modules():
            if isinstance(m, nn.Dropout):
                m.p = 0.0
        g = torch.Generator().manual_seed(11)
        x = torch.randn((n, 3, h, w), generator=g)
        dy = torch.randn((n, classes, h, w), generator=g)
        y = net(x)
        y.backward(dy)
        names, gsum, gabs = [], [], []
        for k, p in net.named_parameters():
            if p.grad is not None:
                names.append(k); gsum.append(float(p.grad.double().sum())); gabs.append(float(p.grad.double().abs().sum()))
        running = {k: v.numpy() for k, v in net.state_dict().items()
                   if 'classifier.project.1.running' in k or 'classifier.2.running' in k or 'final_dec_bn.running' in k}
        np.savez_compressed(os.path.join(OUT, 'net_%s.npz' % tag), x=x.numpy(), dy=dy.numpy(), logits=y.detach().numpy(),
                            grad_names=np.array(names), grad_sum=np.array(gsum), grad_abs=np.array(gabs), **running)


def gen_iteration():
    """Two iterations of the reference loop body with the reference's own classes (DeepLab v2, frozen BN, Adam
    on [pretrained x0.1 (with duplicates), new], EMA 0.99, CutMix var loss, conf_thresh 0.6)."""
    kind, classes, n, h, w, lr = 'resnet101_deeplab_imagenet', 21, 2, 33, 33, 3e-5
    student = build(kind, classes, seed=3, gain=20.0)
    teacher = network_architectures.seg.get(kind)(classes, pretrained=False)
    for p in teacher.parameters():
        p.requires_grad = False
    optim = torch.optim.Adam([dict(params=student.pretrained_parameters(), lr=lr * 0.1),
                              dict(params=student.new_parameters(), lr=lr)], foreach=False)
    ema = optim_weight_ema.EMAWeightOptimizer(teacher, student, 0.99)
    mg = mask_gen.BoxMaskGenerator(prop_range=0.5, n_boxes=1, random_aspect_ratio=True, prop_by_area=True,
                                   within_bounds=True, invert=True)
    crit = nn.CrossEntropyLoss(ignore_index=255)
    rec = dict(kind=kind, classes=classes, n=n, h=h, w=w, lr=lr, conf_thresh=0.6, seed=3, gain=20.0, steps=[])
    for it in range(2):
        g = torch.Generator().manual_seed(100 + it)
        sup_x = torch.randn((n, 3, h, w), generator=g)
        sup_y = torch.randint(0, classes, (n, 1, h, w), generator=g); sup_y[:, :, :4] = 255
        ux0 = torch.randn((n, 3, h, w), generator=g); ux1 = torch.randn((n, 3, h, w), generator=g)
        um0 = torch.ones((n, 1, h, w)); um0[:, :, :3] = 0; um1 = torch.ones((n, 1, h, w)); um1[:, :, :, 16] = 0.5
        masks = torch.tensor(mg.generate_params(n, (h, w), rng=np.random.RandomState(7 + it)).astype(np.float32))
        student.train(); teacher.train(); student.freeze_batchnorm(); teacher.freeze_batchnorm()     # :268-275
        optim.zero_grad()                                                                           # :290
        sup_loss = crit(student(sup_x), sup_y[:, 0]); sup_loss.backward()                            # :299-301
        ux_mixed = ux0 * (1 - masks) + ux1 * masks; um_mixed = um0 * (1 - masks) + um1 * masks       # :350-351
        with torch.no_grad():
            l0 = teacher(ux0).detach(); l1 = teacher(ux1).detach()                                   # :354-356
        ls = student(ux_mixed)                                                                       # :358
        lt = l0 * (1 - masks) + l1 * masks                                                           # :363
        pt = F.softmax(lt, dim=1); ps = F.softmax(ls, dim=1)                                         # :366-367
        conf = (pt.max(dim=1)[0] >= 0.6).float()[:, None]                                            # :409-411
        conf_rate = float(conf.mean())
        loss_mask = um_mixed * conf.mean()                                                           # :415-418
        d = ps - pt
        cons = ((d * d).sum(dim=1, keepdim=True) * loss_mask).mean()                                 # :429-431, :451
        (cons * 1.0).backward()                                                                      # :458-459
        optim.step(); ema.step()                                                                     # :465-467
        tsd, ssd = teacher.state_dict(), student.state_dict()
        rec['steps'].append(dict(sup_loss=float(sup_loss), cons_loss=float(cons), conf_rate=conf_rate,
                                 teacher_abs_sum=float(sum(v.double().abs().sum() for v in tsd.values() if v.dtype == torch.float32)),
                                 student_abs_sum=float(sum(v.double().abs().sum() for v in ssd.values() if v.dtype == torch.float32)),
                                 student_conv1_sum=float(ssd['conv1.weight'].double().sum()),
                                 student_l3_sum=float(ssd['layer3.5.conv2.weight'].double().sum()),
                                 student_l5_sum=float(ssd['layer5.conv2d_list.0.weight'].double().sum()),
                                 teacher_l5_sum=float(tsd['layer5.conv2d_list.1.weight'].double().sum())))
    json.dump(rec, open(os.path.join(OUT, 'iteration.json'), 'w'), indent=1)


def gen_loss_block():
    """SURVEY.md §8c recipe evaluated with the reference formulas (:363-451) -> known answers."""
    torch.manual_seed(0)
    N, C, H, W = 2, 5, 6, 6
    l0 = torch.randn(N, C, H, W) * 4; l1 = torch.randn(N, C, H, W) * 4; ls0 = torch.randn(N, C, H, W) * 4
    um0 = torch.ones(N, 1, H, W); um1 = torch.ones(N, 1, H, W); um0[:, :, 0] = 0; um1[:, :, :, 0] = 0.5
    m = torch.tensor(mask_gen.BoxMaskGenerator(0.5, invert=True).generate_params(N, (H, W), rng=np.random.RandomState(0)).astype(np.float32))
    tau = 0.6
    out = dict(recipe='SURVEY.md 8c', mask_sums=[float(s) for s in m.reshape(N, -1).sum(1)], cases={})
    root_c = C ** 0.5
    for fn in ('var', 'logits_var', 'logits_smoothl1', 'bce', 'kld'):
        for pp in (False, True):
            ls = ls0.clone().requires_grad_(True)
            lt = l0 * (1 - m) + l1 * m
            um = um0 * (1 - m) + um1 * m
            pt = F.softmax(lt, dim=1); ps = F.softmax(ls, dim=1)
            conf = (pt.max(dim=1)[0] >= tau).float()[:, None]
            lm = um * (conf if pp else conf.mean())
            if fn == 'var':
                q = ((ps - pt) ** 2).sum(dim=1, keepdim=True)
            elif fn == 'logits_var':
                q = ((ls - lt) ** 2).sum(dim=1, keepdim=True) / root_c
            elif fn == 'logits_smoothl1':
                q = F.smooth_l1_loss(ls, lt, reduction='none').sum(dim=1, keepdim=True) / root_c
            elif fn == 'bce':
                q = network_architectures.robust_binary_crossentropy(ps, pt).sum(dim=1, keepdim=True)
            else:
                q = F.kl_div(F.log_softmax(ls, dim=1), pt, reduction='none').sum(dim=1, keepdim=True)
            loss = (q * lm).mean(); loss.backward()
            out['cases']['%s_pp%d' % (fn, int(pp))] = dict(loss=float(loss), grad_l1=float(ls.grad.abs().sum()), conf_rate=float(conf.mean()))
    torch.manual_seed(1)
    lg = (torch.randn(2, 5, 6, 6) * 2).requires_grad_(True)
    y = torch.randint(0, 5, (2, 1, 6, 6)); y[:, :, 0] = 255
    ce = nn.CrossEntropyLoss(ignore_index=255)(lg, y[:, 0]); ce.backward()
    out['ce'] = dict(loss=float(ce), grad_l1=float(lg.grad.abs().sum()), n_valid=int((y != 255).sum()))
    out['sigmoid_rampup_3_10'] = network_architectures.sigmoid_rampup(3, 10)
    # EMA known answer on 1,000,003 random elements (reference class on bare modules)
    json.dump(out, open(os.path.join(OUT, 'loss_block.json'), 'w'), indent=1)


def gen_ict_block():
    """ICT loss block: the reference's OWN source lines (train_seg_semisup_ict.py, from the Beta draw to the ramp-up
    multiplication) are cut out of the script and executed on small seeded tensors; teacher_net / student_net are stand-ins
    returning fixed logits.  -> tests/golden/ict_block.json (inputs are re-created from the recipe by the tests)."""
    import textwrap
    lines = open(os.path.join(REF, 'train_seg_semisup_ict.py')).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.strip() == '# ICT mix factors')
    end = next(i for i, l in enumerate(lines) if l.strip() == 'consistency_loss = consistency_loss * ramp_val') + 1
    block = textwrap.dedent('\n'.join(lines[start:end]))
    assert 'ict_mix_factors = np.random.beta' in block and '[:, None, :, :]' in block
    N, C, H, W = 3, 5, 6, 6
    out = dict(recipe='manual_seed(0); l0,l1,ls = randn(3,5,6,6)*4; x0,x1 = randn(3,3,6,6); um0[:,:,0]=0; um1[:,:,:,0]=0.5; '
                      'np.random.seed(case index); ict_alpha=0.4; tau=0.6', ref_lines=[start + 1, end], cases={})
    torch.manual_seed(0)
    l0 = torch.randn(N, C, H, W) * 4; l1 = torch.randn(N, C, H, W) * 4; ls0 = torch.randn(N, C, H, W) * 4
    x0 = torch.randn(N, 3, H, W); x1 = torch.randn(N, 3, H, W)
    um0 = torch.ones(N, 1, H, W); um1 = torch.ones(N, 1, H, W); um0[:, :, 0] = 0; um1[:, :, :, 0] = 0.5
    idx = 0
    for fn in ('var', 'logits_var', 'logits_smoothl1', 'bce', 'kld'):
        for tau, pp, rampup in ((0.6, False, -1), (0.6, True, -1), (0.0, False, 5)):
            ls = ls0.clone().requires_grad_(True)
            seen = {}

            def teacher_net(x):
                return l0 if x is x0 else l1

            def student_net(x):
                seen['mixed'] = x.detach().clone()
                return ls
            np.random.seed(idx)
            ns = dict(np=np, torch=torch, F=F, network_architectures=network_architectures, ict_alpha=0.4,
                      batch_ux0_tea=x0, batch_ux1_tea=x1, batch_ux0_stu=x0, batch_ux1_stu=x1, batch_um0=um0, batch_um1=um1,
                      torch_device=torch.device('cpu'), teacher_net=teacher_net, student_net=student_net,
                      conf_thresh=tau, conf_per_pixel=pp, conf_rate_acc=0.0, rampup=rampup, ramp_val=0.25,
                      cons_loss_fn=fn, root_n_classes=math.sqrt(C))
            exec(block, ns)
            loss = ns['consistency_loss']
            loss.backward()
            out['cases']['%s_tau%g_pp%d_ramp%d' % (fn, tau, int(pp), rampup)] = dict(
                seed=idx, factors=[float(v) for v in ns['ict_mix_factors'].reshape(-1)], loss=float(loss),
                grad_l1=float(ls.grad.abs().sum()), grad_max=float(ls.grad.abs().max()),
                conf_rate_acc=float(ns['conf_rate_acc']), mixed_sum=float(seen['mixed'].double().sum()),
                um_mixed_sum=float(ns['batch_um_mixed'].double().sum()))
            idx += 1
    json.dump(out, open(os.path.join(OUT, 'ict_block.json'), 'w'), indent=1)


def gen_aug_block():
    """Augmentation-consistency loss block: the reference's OWN source lines (train_seg_semisup_aug_mt.py, from the teacher
    prediction to the ramp-up multiplication, :291-394) executed on small seeded tensors with stand-in networks returning
    fixed logits.  `affine_align_corners_kw` is restated as dict(align_corners=True) (datapipe/torch_utils.py:10-12 for
    torch >= 1.3; that module imports distutils, which Python 3.12 no longer ships).  -> tests/golden/aug_block.json, plus
    the three resampled tensors of the block as checksums (inputs are re-created from tests/aug_recipe.py)."""
    import textwrap
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
    from aug_recipe import aug_inputs
    lines = open(os.path.join(REF, 'train_seg_semisup_aug_mt.py')).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.strip() == '# Get teacher predictions for image0')
    end = next(i for i, l in enumerate(lines) if l.strip() == 'consistency_loss = consistency_loss * ramp_val') + 1
    block = textwrap.dedent('\n'.join(lines[start:end]))
    assert 'F.affine_grid(batch_ufx0_to_1' in block and 'F.grid_sample(prob_cons_tea' in block
    lt, ls0, x0, x1, um0, um1, theta = aug_inputs()
    C = lt.shape[1]
    out = dict(recipe='tests/aug_recipe.py::aug_inputs(); tau=0.6; ramp_val=0.25', ref_lines=[start + 1, end], cases={})
    for fn in ('var', 'logits_var', 'logits_smoothl1', 'bce', 'kld'):
        for tau, pp, rampup in ((0.6, False, -1), (0.6, True, -1), (0.0, False, 5)):
            ls = ls0.clone().requires_grad_(True)
            ns = dict(np=np, torch=torch, F=F, network_architectures=network_architectures,
                      affine_align_corners_kw=dict(align_corners=True),
                      batch_ux0=x0, batch_ux1=x1, batch_um0=um0, batch_um1=um1, batch_ufx0_to_1=theta,
                      teacher_net=lambda x: lt, student_net=lambda x: ls,
                      conf_thresh=tau, conf_per_pixel=pp, conf_rate_acc=0.0, rampup=rampup, ramp_val=0.25,
                      cons_loss_fn=fn, root_n_classes=math.sqrt(C))
            key = '%s_tau%g_pp%d_ramp%d' % (fn, tau, int(pp), rampup)
            try:
                exec(block, ns)
            except NameError as e:              # the logits_var branch reads `delta_prob` before assignment (:373)
                out['cases'][key] = dict(raises=type(e).__name__, message=str(e))
                continue
            loss = ns['consistency_loss']
            loss.backward()
            out['cases'][key] = dict(
                loss=float(loss), grad_l1=float(ls.grad.abs().sum()), grad_max=float(ls.grad.abs().max()),
                conf_rate_acc=float(ns['conf_rate_acc']),
                mask_sum=float(ns['mask_tea_in_stu'].double().sum()),
                logits_in_stu_abs_sum=float(ns['logits_cons_tea_in_stu'].double().abs().sum()),
                prob_in_stu_sum=float(ns['prob_cons_tea_in_stu'].double().sum()))
    json.dump(out, open(os.path.join(OUT, 'aug_block.json'), 'w'), indent=1)


def gen_vat_block():
    """VAT perturbation: the reference's OWN helper functions (train_seg_semisup_vat_mt.py:214-301: t_dot, normalize_eps,
    normalized_noise_like, vat_direction, vat_perburbation -- closures of the job function) are cut out of the script and
    executed with a small seeded stand-in for `vat_dir_net`.  -> tests/golden/vat_block.json (inputs: tests/vat_recipe.py)."""
    import textwrap
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
    from vat_recipe import vat_net, vat_inputs
    lines = open(os.path.join(REF, 'train_seg_semisup_vat_mt.py')).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.strip() == 'def t_dot(a, b):')
    end = next(i for i, l in enumerate(lines) if l.strip() == 'return (eps_adv_nrm * adv_radius).detach(), y_pred_logits, y_pred_prob') + 1
    block = textwrap.dedent('\n'.join(lines[start:end]))
    assert 'def vat_direction(x, x_hat):' in block and 'torch.autograd.grad(' in block
    net = vat_net()
    x, x_hat = vat_inputs()
    out = dict(recipe='tests/vat_recipe.py; torch.manual_seed(case seed) before vat_perburbation; vat_radius=0.5',
               ref_lines=[start + 1, end], cases={})
    seed = 100
    for fn in ('var', 'bce', 'kld', 'logits_var'):
        for adaptive in (False, True):
            ns = dict(torch=torch, F=F, math=math, network_architectures=network_architectures, vat_dir_net=net,
                      cons_loss_fn=fn, adaptive_vat_radius=adaptive, vat_radius=0.5)
            exec(block, ns)
            torch.manual_seed(seed)
            x_perturb, y_logits, y_prob = ns['vat_perburbation'](x, x_hat, None)
            assert not x_perturb.requires_grad
            out['cases']['%s_adaptive%d' % (fn, int(adaptive))] = dict(
                seed=seed, abs_sum=float(x_perturb.double().abs().sum()),
                norms=[float(v) for v in x_perturb.reshape(3, -1).double().norm(dim=1)],
                probe=[float(v) for v in x_perturb[:, :, 3, 5].reshape(-1)],
                y_logits_abs_sum=float(y_logits.double().abs().sum()))
            seed += 1
    json.dump(out, open(os.path.join(OUT, 'vat_block.json'), 'w'), indent=1)


def gen_sibling_iterations():
    """Two full iterations of the augmentation-consistency and ICT loops (DeepLab v2) and of the VAT loop (DeepLab v3+, so that the
    eval-mode persistence of the direction network reaches train-mode BatchNorm layers) with the reference's OWN classes
    (networks, EMAWeightOptimizer), torch Adam on the reference's parameter groups, and the unsupervised branch executed from
    the reference scripts' own source lines (train_seg_semisup_aug_mt.py:295-398, train_seg_semisup_vat_mt.py:214-301 +
    397-464, train_seg_semisup_ict.py:305-391).  Dropout probability of the DeepLab v3+ head is set to 0 in both networks (the draw of nn.Dropout cannot be
    shared with another implementation).  -> tests/golden/sibling_iterations.json"""
    import textwrap
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
    from aug_recipe import affine_thetas

    def cut(script, first, last):
        lines = open(os.path.join(REF, script)).read().splitlines()
        i0 = next(i for i, l in enumerate(lines) if l.strip() == first)
        i1 = next(i for i, l in enumerate(lines) if i > i0 and l.strip() == last) + 1
        return textwrap.dedent('\n'.join(lines[i0:i1])), [i0 + 1, i1]

    aug_block, aug_lines = cut('train_seg_semisup_aug_mt.py', '# Get teacher predictions for image0', 'unsup_loss.backward()')
    vat_helpers, vat_hl = cut('train_seg_semisup_vat_mt.py', 'def t_dot(a, b):',
                              'return (eps_adv_nrm * adv_radius).detach(), y_pred_logits, y_pred_prob')
    vat_block, vat_lines = cut('train_seg_semisup_vat_mt.py', '# Compute VAT perburbation', 'unsup_loss.backward()')
    ict_block, ict_lines = cut('train_seg_semisup_ict.py', '# ICT mix factors', 'unsup_loss.backward()')
    out = dict(ref_lines=dict(aug=aug_lines, vat_helpers=vat_hl, vat=vat_lines, ict=ict_lines), runs={})
    crit = nn.CrossEntropyLoss(ignore_index=255)
    for mode, kind, classes, lr, gain in (('aug', 'resnet101_deeplab_imagenet', 21, 3e-5, 4.0),
                                          ('vat', 'resnet101_deeplabv3plus_imagenet', 19, 1e-5, 4.0),
                                          ('ict', 'resnet101_deeplab_imagenet', 21, 3e-5, 4.0),
                                          ('aug_config4', 'densenet161unet', 2, 1e-5, 4.0)):      # BASELINE config 4
        run = mode
        mode = mode.split('_')[0]
        n, h, w = (2, 32, 64) if run == 'aug_config4' else (2, 33, 33)
        student = build(kind, classes, seed=3, gain=gain)
        teacher = (network_architectures.seg.get(kind)(classes) if kind == 'densenet161unet'
                   else network_architectures.seg.get(kind)(classes, pretrained=False))
        for p in teacher.parameters():
            p.requires_grad = False
        for net in (student, teacher):
            for m in net.modules():
                if isinstance(m, nn.Dropout):
                    m.p = 0.0
        optim = torch.optim.Adam([dict(params=student.pretrained_parameters(), lr=lr * 0.1),
                                  dict(params=student.new_parameters(), lr=lr)], foreach=False)
        ema = optim_weight_ema.EMAWeightOptimizer(teacher, student, 0.99)
        rec = dict(kind=kind, classes=classes, n=n, h=h, w=w, lr=lr, seed=3, gain=gain, conf_thresh=0.5, cons_weight=0.7,
                   cons_loss_fn={'aug': 'var', 'vat': 'kld', 'ict': 'bce'}[mode], conf_per_pixel=(mode != 'aug'), vat_radius=0.5, ict_alpha=0.4,
                   adaptive_vat_radius=True, steps=[])
        student.train(); teacher.train(); student.freeze_batchnorm(); teacher.freeze_batchnorm()      # epoch start
        ns = dict(np=np, torch=torch, F=F, math=math, network_architectures=network_architectures,
                  affine_align_corners_kw=dict(align_corners=True), teacher_net=teacher, student_net=student,
                  conf_thresh=0.5, conf_per_pixel=rec['conf_per_pixel'], rampup=-1, ramp_val=1.0, cons_loss_fn=rec['cons_loss_fn'],
                  root_n_classes=math.sqrt(classes), cons_weight=0.7, vat_dir_net=teacher, adaptive_vat_radius=True,
                  vat_radius=0.5, ict_alpha=0.4, torch_device=torch.device('cpu'))
        if mode == 'vat':
            exec(vat_helpers, ns)
        for it in range(2):
            g = torch.Generator().manual_seed(300 + it)
            sup_x = torch.randn((n, 3, h, w), generator=g)
            sup_y = torch.randint(0, classes, (n, 1, h, w), generator=g); sup_y[:, :, :4] = 255
            ux0 = torch.randn((n, 3, h, w), generator=g); ux1 = ux0 + 0.1 * torch.randn((n, 3, h, w), generator=g)
            um0 = torch.ones((n, 1, h, w)); um0[:, :, :3] = 0; um1 = torch.ones((n, 1, h, w)); um1[:, :, :, 16] = 0.5
            optim.zero_grad()
            sup_loss = crit(student(sup_x), sup_y[:, 0]); sup_loss.backward()
            ns['conf_rate_acc'] = 0.0
            if mode == 'aug':
                ns.update(batch_ux0=ux0, batch_ux1=ux1, batch_um0=um0, batch_um1=um1,
                          batch_ufx0_to_1=affine_thetas()[it:it + 2] if it == 0 else affine_thetas()[[2, 0]])
                exec(aug_block, ns)
            elif mode == 'ict':
                np.random.seed(700 + it)                        # the draw of np.random.beta (:306)
                ns.update(batch_ux0_tea=ux0, batch_ux0_stu=ux0, batch_ux1_tea=ux1, batch_ux1_stu=ux1, batch_um0=um0, batch_um1=um1)
                exec(ict_block, ns)
            else:
                torch.manual_seed(500 + it)                     # the draw of normalized_noise_like (:222)
                ns.update(batch_ux_tea=ux0, batch_ux_stu=ux1, batch_um=um0)
                exec(vat_block, ns)
            optim.step(); ema.step()
            tsd, ssd = teacher.state_dict(), student.state_dict()
            rec['steps'].append(dict(
                sup_loss=float(sup_loss), cons_loss=float(ns['consistency_loss']), conf_rate=float(ns['conf_rate_acc']),
                teacher_training=bool(teacher.training), student_training=bool(student.training),
                teacher_abs_sum=float(sum(v.double().abs().sum() for v in tsd.values() if v.dtype == torch.float32)),
                student_abs_sum=float(sum(v.double().abs().sum() for v in ssd.values() if v.dtype == torch.float32)),
                student_conv1_sum=float(ssd[{'vat': 'deeplab.backbone.conv1.weight', 'aug_config4': 'base_model.features.conv0.weight'}
                                            .get(run, 'conv1.weight')].double().sum())))
            if mode == 'ict':
                rec['steps'][-1]['factors'] = [float(v) for v in ns['ict_mix_factors'].reshape(-1)]
        out['runs'][run] = rec
    json.dump(out, open(os.path.join(OUT, 'sibling_iterations.json'), 'w'), indent=1)


def gen_entry_point():
    """Reference CLI / job-function surface of the two drop-in scripts (train_seg_semisup_mask_mt.py:16-42, 581-650;
    train_seg_semisup_ict.py:4-14, 508-577)."""
    import importlib
    import inspect
    import click
    out = dict(scripts={}, rampup=[network_architectures.sigmoid_rampup(e, 10) for e in range(0, 12)])
    for name in ('train_seg_semisup_mask_mt', 'train_seg_semisup_ict', 'train_seg_semisup_aug_mt', 'train_seg_semisup_vat_mt'):
        m = importlib.import_module(name)
        assert os.path.realpath(m.__file__).startswith(os.path.realpath(REF))
        opts = []
        for p in m.experiment.params:
            opts.append(dict(name=p.name, opts=list(p.opts), is_flag=bool(getattr(p, 'is_flag', False)),
                             default=None if callable(p.default) else p.default, type=type(p.type).__name__,
                             choices=list(p.type.choices) if isinstance(p.type, click.Choice) else None))
        job = getattr(m, name)                     # job_helper.job returns the function itself with .submit attached
        out['scripts'][name] = dict(options=opts, job_params=list(inspect.signature(job).parameters),
                                    has_submit=hasattr(job, 'submit'))
    json.dump(out, open(os.path.join(OUT, 'entry_point.json'), 'w'), indent=1)


def gen_input_pipeline():
    """the reference's OWN transform classes SegCVTransformRandomCrop -> SegCVTransformRandomFlip ->
    SegCVTransformNormalizeToTensor (datapipe/seg_transforms_cv.py, imported unmodified; OpenCV is installed) applied to seeded
    uint8 samples of the cases in tests/input_recipe.py, single samples and pairs.  scikit-image is not installed, so the one
    function the module takes from it, `img_as_float`, is provided as `np.multiply(a, 1/255, dtype=float64)` (skimage/util/dtype.py
    `convert` for uint8 input; parity unpinned for that call).  -> tests/golden/input_pipeline.npz (outputs of every case)"""
    import types
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
    import input_recipe as IR
    fake = types.ModuleType('skimage')
    fake.img_as_float = lambda a: np.multiply(a, 1. / 255, dtype=np.float64)
    sys.modules.setdefault('skimage', fake)
    from datapipe import seg_transforms_cv as TCV
    assert os.path.realpath(TCV.__file__).startswith(os.path.realpath(REF))
    out = {}
    for name, case in IR.CASES.items():
        crop = TCV.SegCVTransformRandomCrop(case['crop_size'], case['crop_offset'], rng=np.random.RandomState(case['seed']))
        flip = TCV.SegCVTransformRandomFlip(case['hflip'], case['vflip'], case['hvflip'], rng=np.random.RandomState(case['seed'] + 1))
        norm = TCV.SegCVTransformNormalizeToTensor(None if case['mean'] is None else np.array(case['mean']),
                                                   None if case['std'] is None else np.array(case['std']))
        samples = IR.make_samples(case)
        res = []
        tvt_xf = None
        if case.get('colour'):
            import torchvision.transforms as tvt
            c = case['colour']
            tvt_xf = TCV.SegCVTransformTVT(tvt.Compose([                                     # train_seg_semisup_mask_mt.py:169-179
                tvt.RandomApply([tvt.ColorJitter(c['brightness'], c['contrast'], c['saturation'], c['hue'])], p=c['p']),
                tvt.RandomGrayscale(p=c['grey_p'])]))
            torch.manual_seed(case['torch_seed'])
        for smp in samples:
            if case['pair']:
                a, b = crop.transform_pair(dict(smp), dict(smp))
                a, b = flip.transform_pair(a, b)
                if tvt_xf is not None:
                    a, b = tvt_xf.transform_pair(a, b)
                a, b = norm.transform_pair(a, b)
                res.extend([a, b])
            else:
                a = norm.transform_single(flip.transform_single(crop.transform_single(dict(smp))))
                res.append(a)
        out[name + '.image'] = np.stack([r['image'] for r in res])
        if 'labels' in res[0]:
            out[name + '.labels'] = np.stack([r['labels'] for r in res])
        if 'mask' in res[0]:
            out[name + '.mask'] = np.stack([r['mask'] for r in res])
        print(' ', name, out[name + '.image'].shape, out[name + '.image'].dtype)
    np.savez_compressed(os.path.join(OUT, 'input_pipeline.npz'), **out)


def gen_geom_pipeline():
    """SURVEY.md 8f row 4, scale / rotation crops: the reference's OWN SegCVTransformRandomCropScaleHung / SegCVTransformRandomCropRotateScale
    (datapipe/seg_transforms_cv.py:169-449, imported unmodified; they call cv2.resize / cv2.warpAffine of the installed OpenCV) ->
    SegCVTransformRandomFlip -> SegCVTransformNormalizeToTensor on the seeded uint8 samples of tests/geom_recipe.py, single samples
    and pairs, each sample entering with an identity `xf_cv` so that the transform's matrix is recorded as well (after the crop
    stage, before the flips).  `img_as_float` substituted as in gen_input_pipeline.  -> tests/golden/geom_pipeline.npz"""
    import types
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
    import geom_recipe as GR
    fake = types.ModuleType('skimage')
    fake.img_as_float = lambda a: np.multiply(a, 1. / 255, dtype=np.float64)
    sys.modules.setdefault('skimage', fake)
    from datapipe import seg_transforms_cv as TCV, affine
    assert os.path.realpath(TCV.__file__).startswith(os.path.realpath(REF))
    out = {}
    for name, case in GR.CASES.items():
        rng = np.random.RandomState(case['seed'])
        if case['kind'] == 'hung':
            crop = TCV.SegCVTransformRandomCropScaleHung(case['crop_size'], case['crop_offset'], uniform_scale=case['uniform_scale'], rng=rng)
        else:
            crop = TCV.SegCVTransformRandomCropRotateScale(case['crop_size'], case['crop_offset'], case['rot_mag'], case['max_scale'],
                                                           uniform_scale=case['uniform_scale'],
                                                           constrain_rot_scale=case['constrain_rot_scale'], rng=rng)
        flip = TCV.SegCVTransformRandomFlip(case['hflip'], case['vflip'], case['hvflip'], rng=np.random.RandomState(case['seed'] + 1))
        norm = TCV.SegCVTransformNormalizeToTensor(None if case['mean'] is None else np.array(case['mean']),
                                                   None if case['std'] is None else np.array(case['std']))
        res, xfs = [], []
        for smp in GR.make_samples(case):
            smp = dict(smp, xf_cv=affine.identity_xf(1)[0])
            if case['pair']:
                a, b = crop.transform_pair(dict(smp), dict(smp))
                xfs.extend([a['xf_cv'], b['xf_cv']])
                a, b = norm.transform_pair(*flip.transform_pair(a, b))
                res.extend([a, b])
            else:
                a = crop.transform_single(dict(smp))
                xfs.append(a.pop('xf_cv'))      # (the reference's single-sample flip cannot take a sample that carries xf_cv)
                res.append(norm.transform_single(flip.transform_single(a)))
        out[name + '.image'] = np.stack([r['image'] for r in res])
        out[name + '.xf_cv'] = np.stack(xfs)
        if 'labels' in res[0]:
            out[name + '.labels'] = np.stack([r['labels'] for r in res])
        if 'mask' in res[0]:
            out[name + '.mask'] = np.stack([r['mask'] for r in res])
        print(' ', name, out[name + '.image'].shape, out[name + '.image'].dtype, out[name + '.xf_cv'].dtype)
    np.savez_compressed(os.path.join(OUT, 'geom_pipeline.npz'), **out)


def gen_train_pipeline():
    """the train-time transform lists exactly as train_seg_semisup_mask_mt.py:147-179 assembles them from the
    reference's OWN classes (SegCVTransformRandomCropScaleHung | RandomCropRotateScale | RandomCrop, RandomFlip, SegTransformToPair,
    SegCVTransformTVT, NormalizeToTensor; SegTransformCompose), with seeded generators handed to the crop and flip transforms
    (shared by the supervised and unsupervised list like in the script), applied sample by sample to the seeded uint8 samples of
    tests/pipeline_recipe.py in the order sup_a, unsup, sup_b.  -> tests/golden/train_pipeline.npz"""
    import types
    import torchvision.transforms as tvt
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
    import pipeline_recipe as PR
    fake = types.ModuleType('skimage')
    fake.img_as_float = lambda a: np.multiply(a, 1. / 255, dtype=np.float64)
    sys.modules.setdefault('skimage', fake)
    from datapipe import seg_transforms, seg_transforms_cv
    assert os.path.realpath(seg_transforms_cv.__file__).startswith(os.path.realpath(REF))
    NET_MEAN, NET_STD = np.array(PR.MEAN), np.array(PR.STD)
    out = {}
    for name, case in PR.CASES.items():
        o = PR.options(case)
        crop_size = case['crop_size']
        rng, flip_rng = np.random.RandomState(case['seed']), np.random.RandomState(case['seed'] + 1)
        # ---- train_seg_semisup_mask_mt.py:147-179 (the script's lines; only `rng=` added to the constructors)
        train_transforms = []
        if o['aug_scale_hung']:
            train_transforms.append(seg_transforms_cv.SegCVTransformRandomCropScaleHung(crop_size, (0, 0), uniform_scale=not o['aug_scale_non_uniform'], rng=rng))
        elif o['aug_max_scale'] != 1.0 or o['aug_rot_mag'] != 0.0:
            train_transforms.append(seg_transforms_cv.SegCVTransformRandomCropRotateScale(
                crop_size, (0, 0), rot_mag=o['aug_rot_mag'], max_scale=o['aug_max_scale'], uniform_scale=not o['aug_scale_non_uniform'],
                constrain_rot_scale=True, rng=rng))
        else:
            train_transforms.append(seg_transforms_cv.SegCVTransformRandomCrop(crop_size, (0, 0), rng=rng))
        if o['aug_hflip'] or o['aug_vflip'] or o['aug_hvflip']:
            train_transforms.append(seg_transforms_cv.SegCVTransformRandomFlip(o['aug_hflip'], o['aug_vflip'], o['aug_hvflip'], rng=flip_rng))
        train_unsup_transforms = train_transforms.copy()
        if o['aug_strong_colour']:
            colour_xforms = tvt.Compose([
                tvt.RandomApply([tvt.ColorJitter(o['aug_colour_brightness'], o['aug_colour_contrast'], o['aug_colour_saturation'],
                                                 o['aug_colour_hue'])], p=o['aug_colour_prob']),
                tvt.RandomGrayscale(p=o['aug_colour_greyscale_prob']),
            ])
            train_unsup_transforms.append(seg_transforms.SegTransformToPair())
            train_unsup_transforms.append(seg_transforms_cv.SegCVTransformTVT(colour_xforms))
        train_transforms.append(seg_transforms_cv.SegCVTransformNormalizeToTensor(NET_MEAN, NET_STD))
        train_unsup_transforms.append(seg_transforms_cv.SegCVTransformNormalizeToTensor(NET_MEAN, NET_STD))
        sup = seg_transforms.SegTransformCompose(train_transforms)
        unsup = seg_transforms.SegTransformCompose(train_unsup_transforms)
        # ----
        torch.manual_seed(case['torch_seed'])
        for part, xf in (('sup_a', sup), ('unsup', unsup), ('sup_b', sup)):
            res = [xf.apply(dict(smp)) for smp in PR.make_samples(case, part)]
            key = name + '.' + part
            if part == 'unsup' and o['aug_strong_colour']:
                for m in ('sample0', 'sample1'):
                    out[key + '.' + m + '.image'] = np.stack([r[m]['image'] for r in res])
                    out[key + '.' + m + '.mask'] = np.stack([r[m]['mask'] for r in res])
            elif part == 'unsup':
                out[key + '.image'] = np.stack([r['image'] for r in res]); out[key + '.mask'] = np.stack([r['mask'] for r in res])
            else:
                out[key + '.image'] = np.stack([r['image'] for r in res]); out[key + '.labels'] = np.stack([r['labels'] for r in res])
        print(' ', name, sorted(k for k in out if k.startswith(name + '.unsup')))
    np.savez_compressed(os.path.join(OUT, 'train_pipeline.npz'), **out)


def gen_aug_pipeline():
    """SURVEY.md 8f rows 3-4, BASELINE config 4's data side: the unsupervised transform list of train_seg_semisup_aug_mt.py:126-163
    (`[SegTransformToPair()] + train_transforms` [+ SegCVTransformTVT] + NormalizeToTensor) built from the reference's OWN classes with
    seeded generators, applied to the unsupervised samples of tests/pipeline_recipe.py (each entering with the identity `xf_cv` the
    data set accessor provides, seg_data.py:95-100), then seg_data.SegCollate._compute_xf_0_to_1 on every pair.
    -> tests/golden/aug_pipeline.npz"""
    import types
    import torchvision.transforms as tvt
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
    import pipeline_recipe as PR
    fake = types.ModuleType('skimage')
    fake.img_as_float = lambda a: np.multiply(a, 1. / 255, dtype=np.float64)
    sys.modules.setdefault('skimage', fake)
    from datapipe import seg_transforms, seg_transforms_cv, seg_data, affine
    assert os.path.realpath(seg_data.__file__).startswith(os.path.realpath(REF))
    NET_MEAN, NET_STD = np.array(PR.MEAN), np.array(PR.STD)
    out = {}
    for name, case in PR.AUG_CASES.items():
        o = PR.options(case)
        crop_size, aug_offset_range, aug_free_scale_rot = case['crop_size'], case['aug_offset_range'], case['aug_free_scale_rot']
        rng, flip_rng = np.random.RandomState(case['seed']), np.random.RandomState(case['seed'] + 1)
        # ---- train_seg_semisup_aug_mt.py:126-163 (the script's lines; only `rng=` added to the constructors)
        train_transforms = []
        if o['aug_scale_hung']:
            train_transforms.append(seg_transforms_cv.SegCVTransformRandomCropScaleHung(crop_size, (aug_offset_range, aug_offset_range), uniform_scale=not o['aug_scale_non_uniform'], rng=rng))
        elif o['aug_max_scale'] != 1.0 or o['aug_rot_mag'] != 0.0:
            train_transforms.append(seg_transforms_cv.SegCVTransformRandomCropRotateScale(
                crop_size, (aug_offset_range, aug_offset_range), rot_mag=o['aug_rot_mag'], max_scale=o['aug_max_scale'],
                uniform_scale=not o['aug_scale_non_uniform'], constrain_rot_scale=not aug_free_scale_rot, rng=rng))
        else:
            train_transforms.append(seg_transforms_cv.SegCVTransformRandomCrop(crop_size, (aug_offset_range, aug_offset_range), rng=rng))
        if o['aug_hflip'] or o['aug_vflip'] or o['aug_hvflip']:
            train_transforms.append(seg_transforms_cv.SegCVTransformRandomFlip(o['aug_hflip'], o['aug_vflip'], o['aug_hvflip'], rng=flip_rng))
        train_unsup_transforms = [seg_transforms.SegTransformToPair()] + train_transforms
        if o['aug_strong_colour']:
            colour_xforms = tvt.Compose([
                tvt.RandomApply([tvt.ColorJitter(o['aug_colour_brightness'], o['aug_colour_contrast'], o['aug_colour_saturation'],
                                                 o['aug_colour_hue'])], p=o['aug_colour_prob']),
                tvt.RandomGrayscale(p=o['aug_colour_greyscale_prob']),
            ])
            train_unsup_transforms.append(seg_transforms_cv.SegCVTransformTVT(colour_xforms))
        train_unsup_transforms.append(seg_transforms_cv.SegCVTransformNormalizeToTensor(NET_MEAN, NET_STD))
        unsup = seg_transforms.SegTransformCompose(train_unsup_transforms)
        # ----
        torch.manual_seed(case['torch_seed'])
        res = []
        for smp in PR.make_samples(case, 'unsup'):
            pair = unsup.apply(dict(smp, xf_cv=affine.identity_xf(1)[0]))
            res.append(seg_data.SegCollate._compute_xf_0_to_1(pair))
        for m in ('sample0', 'sample1'):
            out[name + '.' + m + '.image'] = np.stack([r[m]['image'] for r in res])
            out[name + '.' + m + '.mask'] = np.stack([r[m]['mask'] for r in res])
        out[name + '.xf0_to_1_cv'] = np.stack([r['xf0_to_1_cv'] for r in res])
        out[name + '.xf0_to_1'] = np.stack([r['xf0_to_1'] for r in res])
        print(' ', name, out[name + '.sample0.image'].shape, out[name + '.xf0_to_1'].dtype, out[name + '.xf0_to_1_cv'].dtype)
    np.savez_compressed(os.path.join(OUT, 'aug_pipeline.npz'), **out)


def gen_toy2d():
    """BASELINE config 1: the reference's OWN job function `toy2d_train.train_toy2d` (imported unmodified from /root/reference) run
    on the cases of tests/toy2d_recipe.py with torch.manual_seed(TORCH_SEED).  The reference's `toy2d/generate_data.py` cannot be
    imported here (scikit-image, batchup), so the data-set module it asks for is this repository's drop-in (toy2d/generate_data.py:
    same classes / arithmetic, pinned against the reference's committed pickle by tests/test_toy2d.py); everything the job
    function itself does -- network, loaders, perturbation, losses, Adam, EMAWeightOptimizer (the reference's), report -- is
    reference code.  -> tests/golden/toy2d.json"""
    import contextlib
    import importlib.util
    import io
    import tempfile
    repo = os.path.dirname(HERE)
    sys.path.insert(0, os.path.join(repo, 'tests'))
    import toy2d_recipe as T
    pkg_dir = os.path.join(repo, 'toy2d')
    spec = importlib.util.spec_from_file_location('toy2d', os.path.join(pkg_dir, '__init__.py'), submodule_search_locations=[pkg_dir])
    pkg = importlib.util.module_from_spec(spec); sys.modules['toy2d'] = pkg; spec.loader.exec_module(pkg)
    spec2 = importlib.util.spec_from_file_location('toy2d.generate_data', os.path.join(pkg_dir, 'generate_data.py'))
    mod = importlib.util.module_from_spec(spec2); sys.modules['toy2d.generate_data'] = mod; spec2.loader.exec_module(mod)
    pkg.generate_data = mod
    import toy2d_train                                   # the reference's script
    assert os.path.realpath(toy2d_train.__file__).startswith(os.path.realpath(REF))
    assert os.path.realpath(optim_weight_ema.__file__).startswith(os.path.realpath(REF))

    class Cfg(object):
        def __init__(self, run_dir):
            self.run_dir = run_dir
    out = dict(torch_seed=T.TORCH_SEED, cases={})
    with tempfile.TemporaryDirectory() as tmp:
        mask = os.path.join(tmp, 'mask.png')
        T.write_mask_png(mask)
        for name in T.CASES:
            p = T.params(name, mask)
            torch.manual_seed(T.TORCH_SEED)
            buf = io.StringIO()
            run_dir = os.path.join(tmp, name); os.makedirs(run_dir)
            exact = dict(epochs=[])

            def spy(*a, **k):
                """The job's own print: the report lines carry 6 decimals only, so the un-rounded accumulators and the final
                networks are read from the job function's frame at the moment it prints them."""
                import inspect
                loc = inspect.currentframe().f_back.f_locals
                line = ' '.join(str(v) for v in a)
                if line.startswith('Epoch '):
                    exact['epochs'].append([float(loc['batch_sup_loss_accum']), float(loc['batch_conf_mask_sum_accum']),
                                            float(loc['batch_cons_loss_accum'])])
                elif line.startswith('FINAL RESULT'):
                    exact['error_rate'] = float(loc['err_rate'])
                    for tag in ('student_net', 'teacher_net'):
                        net = loc.get(tag)
                        if net is not None:
                            sd = net.state_dict()
                            exact[tag] = dict(abs_sum=float(sum(v.double().abs().sum() for v in sd.values() if v.dtype == torch.float32)),
                                              l_final=[float(v) for v in sd['l_final.weight'].reshape(-1)[:8]])
                buf.write(line + '\n')
            toy2d_train.print = spy
            try:
                toy2d_train.train_toy2d(Cfg(run_dir), **p)
            finally:
                del toy2d_train.print
            epochs, final = T.parse_report(buf.getvalue())
            assert len(epochs) == p['num_epochs'] and final is not None, buf.getvalue()
            out['cases'][name] = dict(epochs=epochs, final_error_pct=final, n_images=len(os.listdir(run_dir)), exact=exact)
            print(' ', name, exact['epochs'], exact['error_rate'])
    json.dump(out, open(os.path.join(OUT, 'toy2d.json'), 'w'), indent=1)


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1:                      # regenerate selected fixtures only: python gen_golden.py toy2d masks ...
        for name in sys.argv[1:]:
            globals()['gen_' + name](); print(name)
        sys.exit(0)
    gen_entry_point(); print('entry point')
    gen_ict_block(); print('ict block')
    gen_aug_block(); print('aug block')
    gen_vat_block(); print('vat block')
    gen_sibling_iterations(); print('sibling iterations')
    gen_masks(); print('masks')
    gen_state_dicts(); print('state dicts')
    gen_loss_block(); print('loss block')
    gen_nets(); print('nets')
    gen_iteration(); print('iteration')

"""Host-logic test of the engine (graph construction + hand-written backward tape) without a GPU.

The product has a single kernel provider (the CUDA extension).  Here it is swapped for tests/_emu_kernels.py —
a torch-CPU double of the kernel INTERFACE — so that gradient bookkeeping (ReLU-gate / residual / BN-scale
folding into dgrad epilogues, concat slices, strided-dgrad phases, parameter-gradient accumulation) is checked
against autograd on the oracle's functional networks.  The numerical parity of the real kernels is covered by the
`-m gpu` tests."""
import os
import sys
from collections import OrderedDict

import pytest
import torch

HERE = os.path.dirname(__file__)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'oracle'))
import torch_oracle as TO  # noqa: E402
from _emu_kernels import EmuKernels  # noqa: E402
from architectures import network_architectures as na  # noqa: E402
from cutmix_semisup_seg_b200 import netbase  # noqa: E402


@pytest.fixture()
def emu():
    saved = netbase.get_kernels
    k = EmuKernels()
    netbase.set_kernels_factory(lambda n_split=1: k)
    yield k
    netbase.set_kernels_factory(saved)


def _run(kind, n, h, w, classes, freeze, seed):
    torch.manual_seed(seed)
    net = na.seg.get(kind)(classes) if kind == 'densenet161unet' else na.seg.get(kind)(classes, pretrained=False)
    sd = TO.synth_state_dict(net.state_dict(), seed=seed)
    net.load_state_dict(sd)
    net.train()
    if freeze:
        net.freeze_batchnorm()
    dm = None
    for m in net.modules():
        if type(m).__name__ == 'B2Dropout':
            if 'unet' in kind:
                dm = (torch.rand(n, h, w, 64) > 0.3).float()           # final_dec_drop acts at the input resolution
            else:
                dm = (torch.rand(n, -(-h // 8), -(-w // 8), 256) > 0.5).float()
            m.inject([dm])
    x = torch.randn(n, 3, h, w)
    y = net(x)
    dy = torch.randn(y.shape)
    y.backward(dy)
    sd64 = OrderedDict((k, v.double().clone() if v.dtype == torch.float32 else v.clone()) for k, v in sd.items())
    for k, p in net.named_parameters():
        if p.requires_grad:
            sd64[k].requires_grad_(True)
    if kind == 'densenet161unet':
        yo = TO.denseunet_forward(sd64, x.double(), backbone_bn_train=not freeze, head_bn_train=True,
                                  dropout_masks=[dm.permute(0, 3, 1, 2).double()])
    elif 'unet' in kind:
        yo = TO.resunet_forward(sd64, x.double(), backbone_bn_train=not freeze, head_bn_train=True,
                                dropout_masks=[dm.permute(0, 3, 1, 2).double()])
    elif 'deeplabv3_' in kind:
        yo = TO.deeplab3_forward(sd64, x.double(), backbone_bn_train=not freeze, head_bn_train=True,
                                 dropout_masks=[dm.permute(0, 3, 1, 2).double()])
    elif 'v3plus' in kind:
        yo = TO.deeplab3plus_forward(sd64, x.double(), backbone_bn_train=not freeze, head_bn_train=True,
                                     dropout_masks=[dm.permute(0, 3, 1, 2).double()])
    else:
        yo = TO.deeplab2_forward(sd64, x.double(), bn_train=not freeze)
    yo.backward(dy.double())
    errs = []
    for k, p in net.named_parameters():
        if not p.requires_grad:
            assert p.grad is None
            continue
        g = sd64[k].grad
        if g is None:
            assert p.grad is None, k
            continue
        assert p.grad is not None, 'missing gradient for ' + k
        assert p.grad.stride() == p.stride()
        errs.append((p.grad.double() - g).abs().max().item() / (g.abs().max().item() + 1e-30))
    lerr = (y.detach().double() - yo.detach()).abs().max().item() / yo.abs().max().item()
    stat = max((v.double() - sd64[k].detach()).abs().max().item() for k, v in net.state_dict().items() if 'running' in k)
    return lerr, sorted(errs), stat, net


def test_deeplab2_frozen_bn_matches_autograd(emu):
    lerr, errs, stat, net = _run('resnet101_deeplab_imagenet', 2, 33, 41, 21, True, seed=1)
    assert lerr < 1e-5
    assert len(errs) == 108 and errs[-1] < 1e-4          # 104 backbone convs + 2 used layer5 convs (w, b)
    assert stat == 0.0                                   # frozen BN must not touch the running statistics
    assert emu.calls.count('conv_fwd') == 106            # layer5.conv2d_list.2/3 are never executed


def test_deeplab3plus_frozen_backbone_train_head(emu):
    lerr, errs, stat, net = _run('resnet101_deeplabv3plus_imagenet', 3, 33, 41, 19, True, seed=2)
    assert lerr < 1e-4
    assert len(errs) == 341
    assert errs[len(errs) // 2] < 1e-3 and errs[-1] < 5e-2      # train-mode BN over tiny maps is ill-conditioned
    assert stat < 1e-4
    nb = {k: int(v) for k, v in net.state_dict().items() if k.endswith('num_batches_tracked')}
    assert nb['deeplab.classifier.project.1.num_batches_tracked'] == 1
    assert nb['deeplab.backbone.bn1.num_batches_tracked'] == 0
    # frozen-BN parameter gradients: nearly every backbone layer gets its column sums from the dgrad epilogue that
    # finished its output gradient; only strided / slice / pooled consumers fall back to the separate reduction
    # (dgamma itself comes from <W, dW>: no pass over the activations, ConvNode._bn_grads_from_wgrad)
    fused, plain = emu.calls.count('bn_eval_param_grad_wdot+stats'), emu.calls.count('bn_eval_param_grad_wdot')
    assert fused == emu.calls.count('conv_dgrad+stats') and fused + plain == 104 and fused >= 90, (fused, plain)
    assert emu.calls.count('bn_eval_param_grad_from_stats') == 0 and emu.calls.count('bn_eval_param_grad') == 0


def test_deeplab3_frozen_backbone_train_head(emu):
    """torchvision's DeepLab v3 in the reference's wrapper (`resnet101_deeplabv3_imagenet`): same backbone and ASPP as v3+,
    DeepLabHead, one x8 resize."""
    lerr, errs, stat, net = _run('resnet101_deeplabv3_imagenet', 3, 33, 41, 21, True, seed=2)
    assert lerr < 1e-4
    assert len(errs) == 335                               # 104 backbone convs + their 2 x 104 BN affines, 23 head tensors
    assert errs[len(errs) // 2] < 1e-3 and errs[-1] < 5e-2
    assert stat < 1e-4
    nb = {k: int(v) for k, v in net.state_dict().items() if k.endswith('num_batches_tracked')}
    assert nb['deeplab.classifier.2.num_batches_tracked'] == 1 and nb['deeplab.classifier.0.project.1.num_batches_tracked'] == 1
    assert nb['deeplab.backbone.bn1.num_batches_tracked'] == 0


def test_resnet50_unet_frozen_encoder_train_decoder(emu):
    """architectures/resunet.py: nearest up-sampling + skip additions, dropout between the last conv and its BatchNorm; every
    trainable tensor but the encoder's unused `fc` receives a gradient."""
    lerr, errs, stat, net = _run('resnet50unet_imagenet', 2, 32, 64, 11, True, seed=2)
    assert lerr < 1e-4
    assert len(errs) == 180 - 2                           # base_model.fc.{weight,bias} are never used (resunet.py:66-92)
    assert errs[len(errs) // 2] < 1e-3 and errs[-1] < 5e-2
    assert stat < 1e-4
    assert net.base_model.fc.weight.grad is None
    for op in ('upsample2x_add', 'upsample2x_bwd'):
        assert emu.calls.count(op) == 5
    assert emu.calls.count('mul_mask') == 2
    nb = {k: int(v) for k, v in net.state_dict().items() if k.endswith('num_batches_tracked')}
    assert nb['final_dec_bn.num_batches_tracked'] == 1 and nb['decoder0.conv_bn.num_batches_tracked'] == 1
    assert nb['base_model.bn1.num_batches_tracked'] == 0


def test_densenet161_unet_frozen_encoder_train_decoder(emu):
    """architectures/denseunet.py (BASELINE config 4): concatenation buffers written slice by slice, stand-alone eval-mode
    BatchNorms on channel prefixes with their gradients accumulated into the buffer's gradient, average-pool transitions."""
    lerr, errs, stat, net = _run('densenet161unet', 2, 32, 64, 2, True, seed=2)
    assert lerr < 1e-5
    assert len(errs) == 503 - 2                           # base_model.classifier.{weight,bias} are never used
    assert errs[-1] < 1e-4
    assert stat < 1e-4
    assert net.base_model.classifier.weight.grad is None
    assert emu.calls.count('scale_channels') == 78 + 3 + 1        # norm1 of 78 dense layers, 3 transition norms, norm5
    assert emu.calls.count('avgpool2x2') == 3 and emu.calls.count('avgpool2x2_bwd') == 3


def test_densenet161_unet_unfrozen_batchnorm(emu):
    """Without --freeze_bn every encoder norm uses batch statistics over its concatenation prefix (train-mode BN kernels on a
    strided slice, gradient accumulated into the buffer's gradient)."""
    lerr, errs, stat, net = _run('densenet161unet', 3, 64, 64, 2, False, seed=5)
    assert lerr < 1e-5 and stat < 1e-5                    # forward and running statistics: exact
    assert len(errs) == 503 - 2
    # gradients through ~160 consecutive train-mode BatchNorms whose statistics come from 12..3072 values are ill-conditioned in
    # fp32 storage (same effect as the DeepLab v3+ head above): median 7e-3 / max 0.3 here, 1.3e-3 / 0.08 at 4 x 128 x 128
    assert errs[len(errs) // 2] < 2e-2 and errs[-1] < 5e-1
    assert stat < 1e-4
    nb = {k: int(v) for k, v in net.state_dict().items() if k.endswith('num_batches_tracked')}
    assert set(nb.values()) == {1}


def test_gradient_accumulation_over_two_backward_passes(emu):
    torch.manual_seed(0)
    net = na.seg.get('resnet101_deeplab_imagenet')(5, pretrained=False)
    net.load_state_dict(TO.synth_state_dict(net.state_dict(), seed=4))
    net.train(); net.freeze_batchnorm()
    x = torch.randn(1, 3, 17, 17)
    dy = torch.randn(1, 5, 17, 17)
    net(x).backward(dy)
    g1 = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
    net(x).backward(dy)
    for k, p in net.named_parameters():
        if p.grad is not None:
            assert torch.allclose(p.grad, 2 * g1[k], rtol=1e-5, atol=1e-7), k


def test_frozen_bn_affine_gradients_accumulate_over_two_passes(emu):
    """dgamma is SET from the accumulated <W, dW> and dbeta totals: two passes must give exactly twice one pass."""
    torch.manual_seed(0)
    net = na.seg.get('resnet101_deeplabv3plus_imagenet')(4, pretrained=False)
    net.load_state_dict(TO.synth_state_dict(net.state_dict(), seed=5))
    net.train(); net.freeze_batchnorm()
    for m in net.modules():
        if hasattr(m, 'p') and hasattr(m, 'next_mask'):
            m.p = 0.0                                   # deterministic head
    x = torch.randn(2, 3, 17, 17)
    dy = torch.randn(2, 4, 17, 17)
    net(x).backward(dy)
    keys = [k for k, p in net.named_parameters() if 'backbone' in k and 'bn' in k and p.grad is not None]
    assert len(keys) >= 200
    g1 = {k: p.grad.clone() for k, p in net.named_parameters() if k in keys}
    net(x).backward(dy)
    sd = dict(net.named_parameters())
    for k in keys:
        assert torch.allclose(sd[k].grad, 2 * g1[k], rtol=2e-4, atol=1e-6 * float(g1[k].abs().max()) + 1e-12), k


def test_eval_mode_forward_has_no_tape(emu):
    net = na.seg.get('resnet101_deeplabv3plus_imagenet')(3, pretrained=False)
    net.eval()
    with torch.no_grad():
        y = net(torch.randn(1, 3, 17, 17))
    assert y.shape == (1, 3, 17, 17) and not y.requires_grad


@pytest.mark.parametrize('kind,classes', [('resnet101_deeplabv3plus_imagenet', 5), ('resnet101_deeplab_imagenet', 4)])
def test_multi_batch_trunk_equals_consecutive_passes(emu, kind, classes):
    """b2_forward_multi / b2_backward_multi (frozen trunk once over [x1 ; x2], head per mini-batch) must give the logits,
    parameter gradients, BatchNorm running statistics and dropout draws of two consecutive forward/backward passes."""
    sizes, h, w = (2, 3), 33, 25
    torch.manual_seed(7)
    xs = [torch.randn(n, 3, h, w) for n in sizes]
    dys = [torch.randn(n, classes, h, w) for n in sizes]
    masks = [(torch.rand(n, -(-h // 8), -(-w // 8), 256) > 0.5).float() for n in sizes]

    def make():
        net = na.seg.get(kind)(classes, pretrained=False)
        net.load_state_dict(TO.synth_state_dict(net.state_dict(), seed=11))
        net.train(); net.freeze_batchnorm()
        for m in net.modules():
            if type(m).__name__ == 'B2Dropout':
                m.inject([m_.clone() for m_ in masks])
        return net

    seq = make()
    ys = []
    for x, dy in zip(xs, dys):
        y = seq(x)
        y.backward(dy)
        ys.append(y.detach())
    multi = make()
    assert multi.b2_trunk_is_batch_invariant()
    n_conv0 = emu.calls.count('conv_fwd')
    with torch.no_grad():                    # (the engine never relies on autograd; the emulation uses torch ops)
        logits, state = multi.b2_forward_multi(xs, record=True)
        n_conv = emu.calls.count('conv_fwd') - n_conv0
        multi.b2_backward_multi(state, dys)
    for a, b in zip(logits, ys):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-5 * float(b.abs().max()))
    sd_s, sd_m = seq.state_dict(), multi.state_dict()
    for k in sd_s:
        if 'running' in k or 'num_batches' in k:
            assert torch.allclose(sd_s[k].double(), sd_m[k].double(), rtol=1e-5, atol=1e-7), k
    ps, pm = dict(seq.named_parameters()), dict(multi.named_parameters())
    worst = 0.0
    for k, p in ps.items():
        if p.grad is None:
            assert pm[k].grad is None, k
            continue
        assert pm[k].grad is not None, 'missing gradient for ' + k
        err = (pm[k].grad - p.grad).abs().max().item() / (p.grad.abs().max().item() + 1e-30)
        worst = max(worst, err)
    assert worst < 2e-3, worst
    if 'v3plus' in kind:
        assert n_conv == 104 + 10            # trunk AND head convolutions once over both mini-batches (BatchNorm per mini-batch)
    else:
        assert n_conv == 106                 # DeepLab v2: the whole network is the trunk


def test_multi_batch_needs_frozen_trunk(emu):
    net = na.seg.get('resnet101_deeplabv3plus_imagenet')(3, pretrained=False)
    net.train()
    with pytest.raises(RuntimeError):
        net.b2_forward_multi([torch.randn(1, 3, 17, 17)] * 2, record=False)


@pytest.mark.parametrize('kind,classes,student', [('resnet101_deeplab_imagenet', 21, False),
                                                  ('resnet101_deeplabv3plus_imagenet', 19, True),
                                                  ('densenet161unet', 2, True)])
def test_input_gradient_only_backward_matches_autograd(emu, kind, classes, student):
    """VAT's direction pass (train_seg_semisup_vat_mt.py:237-268): eval-mode network, d(loss)/d(image) through the stem
    (data-gradient GEMM over the im2col matrix + col2im), no parameter gradient -- also for a network whose parameters
    require gradients (`--vat_dir_from_student`)."""
    torch.manual_seed(3)
    net = na.seg.get(kind)(classes) if kind == 'densenet161unet' else na.seg.get(kind)(classes, pretrained=False)
    sd = TO.synth_state_dict(net.state_dict(), seed=7)
    net.load_state_dict(sd)
    if not student:
        for p in net.parameters():
            p.requires_grad = False
    net.eval()
    hw = (32, 64) if 'unet' in kind else (33, 41)
    x = torch.randn(2, 3, *hw)
    dy = torch.randn(2, classes, *hw)
    with torch.no_grad():          # the engine never uses autograd; the torch-CPU doubles would otherwise record a graph
        logits, state = net.b2_forward(x, record=True, input_grad=True)
        dx = net.b2_backward(state, dy, param_grads=False)
    assert all(p.grad is None for p in net.parameters())
    assert emu.calls.count('col2im') == 1 and emu.calls.count('conv_wgrad') == 0
    sd64 = OrderedDict((k, v.double().clone() if v.dtype == torch.float32 else v.clone()) for k, v in sd.items())
    x64 = x.double().requires_grad_(True)
    if kind == 'densenet161unet':
        yo = TO.denseunet_forward(sd64, x64, backbone_bn_train=False, head_bn_train=False)
    elif 'v3plus' in kind:
        yo = TO.deeplab3plus_forward(sd64, x64, backbone_bn_train=False, head_bn_train=False)
    else:
        yo = TO.deeplab2_forward(sd64, x64, bn_train=False)
    yo.backward(dy.double())
    assert (logits.double() - yo.detach()).abs().max().item() < 1e-4 * yo.abs().max().item()
    assert dx.shape == x.shape
    # fp32 activations against the fp64 oracle: isolated ReLU-gate / max-pool ties flip (DeepLab v2: max 3e-4 of the range),
    # everything else agrees to ~1e-6
    err = (dx.double() - x64.grad).abs()
    assert err.max().item() < 1e-3 * x64.grad.abs().max().item()
    assert err.median().item() < 1e-5 * x64.grad.abs().median().item()
    # a recorded pass without the request keeps the old contract
    with torch.no_grad():
        logits, state = net.b2_forward(x, record=True)
        assert net.b2_backward(state, dy, param_grads=False) is None

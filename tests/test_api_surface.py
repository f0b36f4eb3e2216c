"""Drop-in surface: registry, state_dict layout, parameter groups, helpers, C ABI exports."""
import ctypes
import json
import os
import re

import pytest
import torch

from architectures import network_architectures as na
import optim_weight_ema

HERE = os.path.dirname(__file__)
GOLD = json.load(open(os.path.join(HERE, 'golden', 'state_dicts.json')))


@pytest.mark.parametrize('kind', sorted(GOLD.keys()))
def test_state_dict_identical_to_reference(kind):
    g = GOLD[kind]
    # the reference's densenet161unet factory takes no `pretrained` argument (network_architectures.py:53-55)
    net = na.seg.get(kind)(g['classes']) if kind == 'densenet161unet' else na.seg.get(kind)(g['classes'], pretrained=False)
    mine = [[k, list(v.shape), str(v.dtype).replace('torch.', '')] for k, v in net.state_dict().items()]
    assert mine == g['entries']
    assert len(list(net.pretrained_parameters())) == g['n_pretrained']
    assert len(set(id(p) for p in net.pretrained_parameters())) == g['n_pretrained_unique']
    assert len(list(net.new_parameters())) == g['n_new']
    assert [k for k, p in net.named_parameters() if p.requires_grad] == g['trainable']
    assert net.BLOCK_SIZE == ((32, 32) if 'unet' in kind else (1, 1))
    if kind == 'densenet161unet':            # no ImageNet statistics without ImageNet weights (denseunet.py:146-148)
        assert net.MEAN is None and net.STD is None
    else:
        assert len(net.MEAN) == 3 and len(net.STD) == 3


def test_registry_names_and_unbuilt_architectures():
    g = GOLD['resnet101_deeplab_imagenet']
    assert sorted(na.seg.names()) == g['registry_names']
    with pytest.raises(NotImplementedError):
        na.seg.get('resnet101_pspnet_imagenet')(21)        # needs the external mit_semseg package, like the reference


def test_freeze_batchnorm_scope():
    v2 = na.seg.get('resnet101_deeplab_imagenet')(21, pretrained=False)
    v2.train(); v2.freeze_batchnorm()
    assert all(not m.training for m in v2.modules() if 'BatchNorm' in type(m).__name__)
    v3 = na.seg.get('resnet101_deeplabv3plus_imagenet')(19, pretrained=False)
    v3.train(); v3.freeze_batchnorm()
    assert all(not m.training for m in v3.deeplab.backbone.modules() if 'BatchNorm' in type(m).__name__)
    assert all(m.training for m in v3.deeplab.classifier.modules() if 'BatchNorm' in type(m).__name__)


def test_ema_optimizer_contract_on_cpu_modules():
    a = torch.nn.Sequential(torch.nn.Linear(3, 2), torch.nn.BatchNorm1d(2))
    b = torch.nn.Sequential(torch.nn.Linear(3, 2), torch.nn.BatchNorm1d(2))
    opt = optim_weight_ema.EMAWeightOptimizer(a, b, 0.99)
    assert len(opt.target_params) == len(opt.source_params) == 6           # weights, biases, running stats
    for t, s in zip(a.state_dict().values(), b.state_dict().values()):
        assert torch.equal(t, s)
    with pytest.raises(ValueError):
        optim_weight_ema.EMAWeightOptimizer(a, torch.nn.Sequential(torch.nn.Linear(3, 2)), 0.99)
    with pytest.raises(RuntimeError):
        opt.step()            # CPU tensors: the B200 path has no CPU fallback and must say so


def test_cpu_input_fails_loudly():
    net = na.seg.get('resnet101_deeplab_imagenet')(21, pretrained=False)
    with pytest.raises(Exception):
        net(torch.zeros(1, 3, 33, 33))


def test_c_abi_exports_every_declared_symbol():
    from cutmix_semisup_seg_b200 import lib
    header = open(os.path.join(os.path.dirname(HERE), 'include', 'b200seg.h')).read()
    declared = set(re.findall(r'\b(b2_[a-z0-9_]+)\s*\(', header))
    declared -= {'b2_ema_chunk', 'b2_conv_params', 'b2_wgrad_params'}
    dll = ctypes.CDLL(lib.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(dll, s)]
    assert not missing, missing
    assert set(lib.exported_symbols()) | {'b2_debug_set', 'b2_debug_trace'} >= declared
    l = lib.load()
    assert l.b2_version() >= 1
    assert l.b2_num_sms() == 0 or l.b2_num_sms() > 0     # no compute call without a GPU


def test_c_abi_rejects_bad_arguments_before_any_launch():
    """Error behaviour of the ABI (include/b200seg.h: non-zero code + b2_last_error(), no exception, no exit): argument
    validation happens before the first CUDA call, so it can be exercised without a GPU.  Pointers below are never
    dereferenced (every call fails validation)."""
    from cutmix_semisup_seg_b200 import lib as L
    fake = 0x1000
    with pytest.raises(L.B2Error, match='bad args'):
        L.call('b2_aug_consistency_fwd_bwd', None, fake, fake, fake, fake, fake, fake, 1, 2, 4, 4, 0, 0.5, 0, None)
    with pytest.raises(L.B2Error, match='unknown loss_fn'):
        L.call('b2_aug_consistency_fwd_bwd', fake, fake, fake, fake, fake, fake, fake, 1, 2, 4, 4, 9, 0.5, 0, None)
    with pytest.raises(L.B2Error, match='unsupported'):
        L.call('b2_aug_consistency_fwd_bwd', fake, fake, fake, fake, fake, fake, fake, 1, 65, 4, 4, 0, 0.5, 0, None)
    with pytest.raises(L.B2Error, match='bad args'):
        L.call('b2_affine_grid_sample', fake, None, fake, 1, 2, 4, 4, 4, 4, None)
    with pytest.raises(L.B2Error, match='l1 given without mix mask'):
        L.call('b2_consistency_fwd_bwd', fake, fake, fake, None, None, fake, fake, 1, 2, 16, 0, 0.5, 0, None)
    with pytest.raises(L.B2Error, match='needs confbar'):
        L.call('b2_ict_consistency_fwd_bwd', fake, fake, fake, fake, None, None, fake, fake, 1, 2, 16, 0, 0.5, 1, None)
    with pytest.raises(L.B2Error, match='bad args'):
        L.call('b2_ce_fwd_bwd', fake, None, fake, fake, 1, 2, 16, 255, None)
    assert L.call('b2_consistency_num_partials', 3, 1000) == 3 * 4          # ceil(1000 / 256) blocks per image

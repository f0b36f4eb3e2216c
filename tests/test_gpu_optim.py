"""-m gpu: fused optimiser + EMA kernel (csrc/optim.cu, cutmix_semisup_seg_b200/optim.py) against torch.optim's per-tensor
implementations (`foreach=False`: what the reference ran, and the only torch path that honours duplicated entries)."""
import copy
import warnings

import pytest
import torch

import optim_weight_ema
from cutmix_semisup_seg_b200.optim import FusedOptimizer

pytestmark = pytest.mark.gpu
dev = torch.device('cuda:0')


class _Net(torch.nn.Module):
    """Parameters of assorted sizes / layouts plus a float buffer and an int64 buffer (EMA covers float state only)."""

    def __init__(self, seed):
        super(_Net, self).__init__()
        g = torch.Generator().manual_seed(seed)
        self.a = torch.nn.Parameter(torch.randn(64, 32, 3, 3, generator=g).contiguous(memory_format=torch.channels_last))
        self.b = torch.nn.Parameter(torch.randn(8193, generator=g))
        self.c = torch.nn.Parameter(torch.randn(1, generator=g))
        self.d = torch.nn.Parameter(torch.randn(19, 256, 1, 1, generator=g))
        self.e = torch.nn.Parameter(torch.randn(7, generator=g), requires_grad=False)       # frozen: EMA only
        self.register_buffer('running', torch.randn(33, generator=g))
        self.register_buffer('count', torch.tensor(3, dtype=torch.long))


def _groups(net, lr):
    # the first group repeats tensors like DeepLab v2's pretrained_parameters(): a x3, b x2, c x1
    return [dict(params=[net.a, net.b, net.a, net.c, net.a, net.b], lr=lr * 0.1), dict(params=[net.d], lr=lr)]


def _set_grads(nets, step):
    g = torch.Generator().manual_seed(100 + step)
    for name in ('a', 'b', 'c', 'd'):
        ref = getattr(nets[0], name)
        gr = torch.randn(ref.shape, generator=g) * (10.0 ** (step - 1))
        for net in nets:
            p = getattr(net, name)
            if p.grad is None:
                p.grad = torch.empty_like(p)            # same (possibly channels-last) layout as the parameter
            p.grad.copy_(gr.to(p.device))


@pytest.mark.parametrize('opt_type', ['adam', 'sgd', 'sgd_nesterov'])
def test_fused_optimizer_matches_torch_per_tensor_loop_with_duplicates(opt_type):
    lr = 1e-3
    ref_net, net, teacher = _Net(1).to(dev), _Net(1).to(dev), _Net(2).to(dev)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        if opt_type == 'adam':
            ref = torch.optim.Adam(_groups(ref_net, lr), foreach=False)
            fused = FusedOptimizer(_groups(net, lr), 'adam', lr=lr)
        else:
            nest = opt_type == 'sgd_nesterov'
            ref = torch.optim.SGD(_groups(ref_net, lr), momentum=0.9, nesterov=nest, weight_decay=5e-4, foreach=False)
            fused = FusedOptimizer(_groups(net, lr), 'sgd', lr=lr, momentum=0.9, nesterov=nest, weight_decay=5e-4)
    ema = optim_weight_ema.EMAWeightOptimizer(teacher, net, 0.99)
    assert 'k sequential' in fused.b2_note
    start = {k: v.detach().clone() for k, v in net.state_dict().items()}
    for step in range(4):
        for grp_r, grp_f in zip(ref.param_groups, fused.param_groups):         # a schedule changes lr every iteration
            grp_r['lr'] = grp_f['lr'] = grp_r['lr'] * 0.9
        _set_grads([ref_net, net], step)
        t_before = {k: v.detach().clone() for k, v in teacher.state_dict().items()}
        ref.step()
        fused.step(ema=ema)
        torch.cuda.synchronize()
        # EMA: bit-exact on the values the kernel itself produced (optim_weight_ema.py:21-25 arithmetic)
        for k, s in net.state_dict().items():
            t = teacher.state_dict()[k]
            if s.dtype == torch.float32:
                want = t_before[k] * 0.99
                want = want + s * (1.0 - 0.99)
                assert torch.equal(t, want), k
            else:
                assert torch.equal(t, t_before[k]), k
    for name in ('a', 'b', 'c', 'd'):
        got, want, p0 = getattr(net, name).detach(), getattr(ref_net, name).detach(), start[name]
        assert got.stride() == want.stride()
        upd = (want - p0).abs().max().item()
        assert upd > 0
        assert (got - want).abs().max().item() <= 2e-5 * upd + 1e-7 * want.abs().max().item(), name
    assert torch.equal(net.e, start['e'])                                   # frozen parameter untouched by the optimiser
    assert int(fused.state[net.a]['b2_steps_done']) == 4
    st = ref.state[ref_net.a]
    if opt_type == 'adam':
        assert int(st['step']) == 12                                        # listed 3 times: 3 updates per step()
        want_m = st['exp_avg'].permute(0, 2, 3, 1).reshape(-1)            # the fused state is kept in memory order
        assert (fused.state[net.a]['exp_avg'].view(-1) - want_m).abs().max().item() <= 1e-5 * want_m.abs().max().item()


def test_fused_optimizer_state_dict_roundtrip_and_plain_step():
    net = _Net(3).to(dev)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        opt = FusedOptimizer(_groups(net, 1e-3), 'adam', lr=1e-3)
    _set_grads([net], 0)
    opt.step()                                  # no EMA partner
    sd = copy.deepcopy(opt.state_dict())
    p_mid = {k: v.detach().clone() for k, v in net.named_parameters()}
    _set_grads([net], 1)
    opt.step()
    p_end = {k: v.detach().clone() for k, v in net.named_parameters()}
    with torch.no_grad():
        for k, v in net.named_parameters():
            v.copy_(p_mid[k])
    opt.load_state_dict(sd)
    assert int(opt.state[net.a]['b2_steps_done']) == 1
    opt.step()
    for k, v in net.named_parameters():
        assert torch.equal(v.detach(), p_end[k]), k


def test_fused_optimizer_rejects_cpu_parameters():
    net = _Net(4)
    with pytest.raises(RuntimeError):
        FusedOptimizer(_groups(net, 1e-3), 'adam', lr=1e-3)

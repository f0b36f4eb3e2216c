"""CPU dry run of cutmix_semisup_seg_b200.kernels.ActKernels (the launch-shape layer between the engine and the C ABI): a
recording backend whose methods BIND their arguments against the real CudaBackend signatures, so that a wrong keyword or a
duplicated argument in the Python glue fails here and not on the GPU box.  Covers the chunked parity-mode launches (KCHUNK)."""
import inspect

import pytest
import torch

from cutmix_semisup_seg_b200 import ops as O
from cutmix_semisup_seg_b200.acts import Act
from cutmix_semisup_seg_b200.kernels import ActKernels


class Recorder(object):
    """Stands in for CudaBackend: every call is bound against the real method's signature and logged."""

    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        real = getattr(O.CudaBackend, name)
        sig = inspect.signature(real)

        def call(*args, **kwargs):
            bound = sig.bind(self, *args, **kwargs)
            bound.apply_defaults()
            self.calls.append((name, bound.arguments))
            if name == 'split_tf32':
                return args[0].clone(), args[0].clone()
            if name == 'transpose_w':
                a = bound.arguments
                return torch.zeros((a['b'], a['t'], a['ldd'] if a['ldd'] is not None else a['a']))
            if name == 'conv_gemm' and bound.arguments['want_stats']:
                assert bound.arguments['device'] is not None
                return (torch.zeros(1), 1, 4)
            return None
        return call


def _acts(n, h, w, cin, cout):
    x = Act(torch.zeros(n, h, w, cin), n, h, w, cin)
    y = Act(torch.zeros(n, h, w, (cout + 3) // 4 * 4), n, h, w, cout, (cout + 3) // 4 * 4)
    return x, y


@pytest.mark.parametrize('n_split', [1, 3])
@pytest.mark.parametrize('cin,cout,k,dil', [(2048, 256, 3, 12), (256, 256, 3, 2), (256, 1024, 1, 1), (2048, 512, 1, 1), (304, 256, 3, 1),
                                            (64, 64, 1, 1)])
def test_conv_fwd_dgrad_wgrad_launch_plumbing(n_split, cin, cout, k, dil):
    rec = Recorder()
    K = ActKernels(backend=rec, n_split=n_split)
    pad = dil * (k // 2)
    x, y = _acts(1, 8, 8, cin, cout)
    w = torch.zeros(cout, k * k, cin)
    res = Act(torch.zeros(1, 8, 8, y.ld), 1, 8, 8, cout, y.ld)
    K.conv_fwd(x, w, cout, k, k, cin, cin, 1, pad, dil, y, scale=torch.ones(cout), shift=torch.zeros(cout), addend=res, relu=True)
    gemms = [c for c in rec.calls if c[0] == 'conv_gemm']
    red = cin * k * k
    if n_split == 1 or red <= K.KCHUNK:
        assert len(gemms) == 1
    else:
        assert len(gemms) == len(K._chunks(O.conv_taps(k, k, dil, pad), cin)) > 1
        # partial launches: scaled, no shift / ReLU, the residual rides on the first one, later ones accumulate
        assert gemms[0][1]['addend'] == res.ptr and not gemms[0][1]['accumulate'] and gemms[0][1]['shift'] is None
        assert all(g[1]['accumulate'] and g[1]['addend'] is None for g in gemms[1:-1])
        assert all(not g[1]['relu'] and g[1]['scale'] is not None for g in gemms[:-1])
        last = gemms[-1][1]
        assert last['relu'] and last['shift'] is not None and last['addend'] not in (None, res.ptr) and not last['accumulate']
        # every reduction term is covered exactly once
        assert sum(g[1]['k'] * len(g[1]['taps']) for g in gemms) == red
    # dgrad with fused addend + gate + statistics
    rec.calls.clear()
    g = Act(torch.zeros(1, 8, 8, y.ld), 1, 8, 8, cout, y.ld)
    dx = Act(torch.zeros(1, 8, 8, cin), 1, 8, 8, cin)
    wt, ldb = K.transpose_w(w, cout, k * k, cin)
    st = K.conv_dgrad(g, wt, cin, k, k, cout, ldb, 1, pad, dil, dx, addend=Act(torch.zeros(1, 8, 8, cin), 1, 8, 8, cin),
                      gate=Act(torch.zeros(1, 8, 8, cin), 1, 8, 8, cin), want_stats=True)
    assert st is not None
    gemms = [c for c in rec.calls if c[0] == 'conv_gemm']
    assert gemms[-1][1]['want_stats'] and gemms[-1][1]['gate'] is not None
    assert all(not g_[1]['want_stats'] and g_[1]['gate'] is None for g_ in gemms[:-1])
    # accumulate into an existing gradient, strided dgrad, wgrad
    rec.calls.clear()
    K.conv_dgrad(g, wt, cin, k, k, cout, ldb, 1, pad, dil, dx, accumulate=True)
    assert [c for c in rec.calls if c[0] == 'conv_gemm'][-1][1]['accumulate']
    K.conv_wgrad(g, x, torch.zeros(cout, k * k, cin), cout, k, k, cin, 1, pad, dil, row_scale=torch.ones(cout), accumulate=True)
    assert any(c[0] == 'conv_wgrad' for c in rec.calls)

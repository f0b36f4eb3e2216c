"""Micro-benchmark of the HBM-bound operators at the cfg3 shapes (DeepLab v3+, N = 16, 512 x 512): CUDA-event time per
call (L2 flushed between calls) and achieved GB/s against the algorithmic bytes of each op.
    python tools/netops_bench.py [reps] [filter]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from cutmix_semisup_seg_b200.kernels import ActKernels  # noqa: E402
from cutmix_semisup_seg_b200.acts import Act  # noqa: E402

dev = torch.device('cuda:0')
K = ActKernels(n_split=1)
be = K.be
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
filt = sys.argv[2] if len(sys.argv) > 2 else ''
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)


def timeit(name, nbytes, fn):
    if filt and filt not in name:
        return
    ts = []
    for i in range(reps + 1):
        flush.fill_(float(i))
        torch.cuda._sleep(400000)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i > 0:
            ts.append(e0.elapsed_time(e1))
    if not ts:
        return                                     # reps == 0: single untimed call (ncu captures)
    ms = sorted(ts)[len(ts) // 2]
    print('{:<46s} {:8.3f} ms  {:8.1f} GB/s  ({:.0f} MB)'.format(name, ms, nbytes / ms / 1e6, nbytes / 1e6), flush=True)


def act(n, h, w, c, ld=None):
    a = Act.alloc(n, h, w, c, dev, ld=ld)
    a.base.normal_()
    return a


N = 16
# stem im2col
x = act(N, 512, 512, 3, ld=4)
timeit('im2col 7x7s2 3->K160 @512', N * 256 * 256 * 160 * 4 + N * 512 * 512 * 16, lambda: K.im2col(x, 7, 7, 2, 3, 1, 256, 256, 160))
# max pool
x = act(N, 256, 256, 64); y = act(N, 128, 128, 64); idx = torch.empty((N, 128, 128, 64), dtype=torch.uint8, device=dev)
timeit('maxpool fwd 64ch 256->128', (x.base.numel() + y.base.numel()) * 4 + idx.numel(), lambda: K.maxpool_fwd(x, y, idx))
dx = act(N, 256, 256, 64)
timeit('maxpool bwd', (x.base.numel() + y.base.numel()) * 4 + idx.numel(), lambda: K.maxpool_bwd(y, idx, dx))
# bilinear NHWC (decoder x2 into the concat slice)
a = act(N, 64, 64, 256); cat = act(N, 128, 128, 304); sl = cat.slice(48, 256)
timeit('bilinear fwd nhwc 256ch 64->128', (a.base.numel() + N * 128 * 128 * 256) * 4, lambda: K.bilinear_fwd(a, sl, False))
timeit('bilinear bwd nhwc 256ch 128->64', (a.base.numel() + N * 128 * 128 * 256) * 4, lambda: K.bilinear_bwd(sl, a, False))
# final resize to NCHW logits
low = act(N, 128, 128, 19, ld=20); logits = torch.empty((N, 19, 512, 512), device=dev)
timeit('bilinear fwd nchw 19ch 128->512', (low.base.numel() + logits.numel()) * 4, lambda: K.bilinear_fwd_nchw(low, logits, False))
sc = torch.ones(1, device=dev)
timeit('bilinear bwd nchw 19ch 512->128', (low.base.numel() + logits.numel()) * 4,
       lambda: K.bilinear_bwd_nchw(logits, low, False, scale_dev=sc))
# train-mode BN (decoder 128x128x256 and ASPP 64x64x256)
for hw, tag in ((128, 'decoder'), (64, 'aspp')):
    x = act(N, hw, hw, 256); y = act(N, hw, hw, 256); dy = act(N, hw, hw, 256); dxa = act(N, hw, hw, 256)
    m = torch.zeros(256, device=dev); r = torch.ones(256, device=dev); g = torch.ones(256, device=dev); b = torch.zeros(256, device=dev)
    rm = torch.zeros(256, device=dev); rv = torch.ones(256, device=dev); dg = torch.zeros(256, device=dev); db = torch.zeros(256, device=dev)
    nb = x.base.numel() * 4
    timeit('bn_stats %s' % tag, nb, lambda: K.bn_stats(x, 1e-5, 0.1, m, r, rm, rv))
    timeit('bn_apply+relu %s' % tag, 2 * nb, lambda: K.bn_apply(x, m, r, g, b, True, None, 1.0, y))
    timeit('bn_bwd (reduce + dx) %s' % tag, 7 * nb, lambda: K.bn_bwd(dy, x, y, m, r, g, True, None, 1.0, dxa, dg, db, False))
# losses
l0 = torch.randn((N, 19, 512, 512), device=dev); l1 = torch.randn_like(l0); ls = torch.randn_like(l0)
mk = (torch.rand((N, 1, 512, 512), device=dev) > 0.5).float(); um = torch.ones_like(mk)
dls = torch.empty_like(ls)
timeit('consistency var 19ch 512^2', 4 * l0.numel() * 4 + 2 * mk.numel() * 4,
       lambda: be.consistency(l0, l1, ls, mk, um, 'var', 0.97, False, 1.0, 1.0, dls=dls))
l0p, l1p, lsp = l0 * 40.0, l1 * 40.0, ls * 40.0          # peaked logits (the bench's conditioned classifier): p == 0 for most classes
timeit('consistency var 19ch 512^2 peaked x40', 4 * l0.numel() * 4 + 2 * mk.numel() * 4,
       lambda: be.consistency(l0p, l1p, lsp, mk, um, 'var', 0.97, False, 1.0, 1.0, dls=dls))
del l0p, l1p, lsp
lab = torch.randint(0, 19, (N, 512, 512), device=dev)
timeit('cross entropy 19ch 512^2', 2 * l0.numel() * 4 + lab.numel() * 8, lambda: be.cross_entropy(l0, lab, dlogits=dls))
# gate / copies / pooling
g1 = act(N, 128, 128, 256); y1 = act(N, 128, 128, 256)
timeit('relu_gate 256ch 128^2', 3 * g1.base.numel() * 4, lambda: K.relu_gate(g1, y1))
timeit('copy_act 256ch 128^2', 2 * g1.base.numel() * 4, lambda: K.copy_act(g1, y1))
x = act(N, 64, 64, 2048); v = Act.alloc(N, 1, 1, 2048, dev, ld=2048)
timeit('gap fwd 2048ch 64^2', x.base.numel() * 4, lambda: K.gap_fwd(x, v))
timeit('gap bwd 2048ch 64^2', x.base.numel() * 4, lambda: K.gap_bwd(v, x))
timeit('colsum 2048ch 64^2', x.base.numel() * 4, lambda: K.colsum(x, torch.empty(2048, device=dev), False))


# frozen-BN parameter gradients from <W, dW> + the epilogue's partial column sums (layer3 conv3: 1024 channels)
class _BN(object):
    pass


bn = _BN(); bn.eps = 1e-5
bn.weight = torch.ones(1024, device=dev); bn.running_mean = torch.zeros(1024, device=dev); bn.running_var = torch.ones(1024, device=dev)
W = torch.randn(1024, 256, device=dev); gW = torch.randn(1024, 256, device=dev)
stats = (torch.randn(2048, 2, 1024, device=dev), 2048, 1024)
dgam = torch.zeros(1024, device=dev); dbet = torch.zeros(1024, device=dev)
gact = act(N, 64, 64, 1024)
timeit('bn wdot from epilogue stats c1024', stats[0].numel() * 4 + 2 * W.numel() * 4,
       lambda: K.bn_eval_param_grad_wdot(stats, gact, W, gW, bn, dgam, dbet, True))
timeit('bn wdot from g (colsum) c1024', gact.base.numel() * 4, lambda: K.bn_eval_param_grad_wdot(None, gact, W, gW, bn, dgam, dbet, True))
w3 = torch.randn(256, 9, 256, device=dev)
timeit('transpose_w 256x9x256', 2 * w3.numel() * 4, lambda: K.transpose_w(w3, 256, 9, 256))
# CutMix image / valid-mask mix (X1), box-mask rasterisation (M2), fused multi-tensor EMA over the DeepLab v3+ state (E1)
import numpy as np  # noqa: E402
x0 = torch.randn((N, 3, 512, 512), device=dev); x1 = torch.randn_like(x0)
timeit('mix image 3ch 512^2', 3 * x0.numel() * 4 + mk.numel() * 4, lambda: be.mix(x0, x1, mk))
timeit('mix valid-mask 1ch 512^2', 4 * mk.numel() * 4, lambda: be.mix(um, um, mk))
boxes = torch.from_numpy(np.tile(np.array([[[100, 120, 400, 380]]], dtype=np.int32), (N, 1, 1))).to(dev)
timeit('box mask rasterise 512^2', mk.numel() * 4, lambda: be.box_mask_rasterize(boxes, 512, 512, 0.0))
n_state = 59453331                      # float32 state elements of DeepLab v3+
t_flat = torch.randn(n_state, device=dev); s_flat = torch.randn(n_state, device=dev)
timeit('EMA flat 59.45M elems', 12 * n_state, lambda: be.ema_step_flat(t_flat, s_flat, 0.99))

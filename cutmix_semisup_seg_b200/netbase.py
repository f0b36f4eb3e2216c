"""Base class of the B200 segmentation networks: glue between nn.Module semantics (parameters,
state_dict, train/eval, autograd) and the engine's explicit forward/backward tape."""
import os

import torch
import torch.nn as nn

from . import engine as E
from .acts import Act

_KERNELS = {}


def get_kernels(n_split=1):
    """Process-wide ActKernels per precision mode (kept out of the modules so they stay picklable)."""
    from .kernels import ActKernels
    if n_split not in _KERNELS:
        _KERNELS[n_split] = ActKernels(n_split=n_split)
    return _KERNELS[n_split]


def set_kernels_factory(fn):
    """Test hook: replace the kernel provider (tests/_emu_kernels.py exercises the engine on CPU)."""
    global get_kernels
    get_kernels = fn


class _RunState(object):
    """What the backward pass needs from one recorded forward pass."""
    __slots__ = ('tape', 'low', 'align', 'consumed', 'xin')

    def __init__(self, tape, low, align, xin=None):
        self.tape, self.low, self.align, self.consumed = tape, low, align, False
        self.xin = xin          # the image Act when its gradient was requested (b2_forward(input_grad=True)), else None


class _MultiRunState(object):
    """Recorded multi-batch pass (b2_forward_multi): one tape, one (low-res logits, align_corners) per mini-batch."""
    __slots__ = ('tape', 'lows', 'consumed')

    def __init__(self, tape, lows):
        self.tape, self.lows, self.consumed = tape, lows, False


class _B2Function(torch.autograd.Function):
    """Bridges torch.autograd to the engine tape for drop-in use (`loss.backward()` on any loss built
    from the returned logits).  Parameter gradients are written straight into `p.grad`."""

    @staticmethod
    def forward(ctx, x, anchor, net):
        logits, state = net.b2_forward(x, record=True)
        ctx.net, ctx.state = net, state
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        ctx.net.b2_backward(ctx.state, dlogits.contiguous())
        return None, None, None


class B2SegNet(nn.Module):
    BLOCK_SIZE = (1, 1)
    # multi-batch passes run the head's convolutions once over all mini-batches too (train-mode BatchNorm per mini-batch);
    # B200SEG_BATCH_HEAD=0 restores the head-per-mini-batch schedule for A/B timing
    b2_batched_head = os.environ.get('B200SEG_BATCH_HEAD', '1') != '0'

    # precision of the tensor-core convolutions: 'tf32' (one pass; what cuDNN does by default) or
    # '3xtf32' (operands split in hi/lo parts, three passes: ~fp32 accuracy, used by the parity tests)
    def __init__(self):
        super(B2SegNet, self).__init__()
        self.b2_precision = os.environ.get('B200SEG_PRECISION', 'tf32')

    def _n_split(self):
        return {'tf32': 1, '3xtf32': 3, '4xtf32': 4}[self.b2_precision]

    # ---- subclasses define the graph --------------------------------------------------------
    # The graph is given in two parts: a TRUNK (the ResNet-101 backbone; DeepLab v2: the whole network) and a HEAD.  With
    # frozen BatchNorm the trunk is batch-invariant (every sample is processed independently), so the training step
    # runs it ONCE over the concatenation of several mini-batches (b2_forward_multi) -- bigger GEMMs, half the
    # launches, no partial last wave on the 256-channel layers.  The head's convolutions run over the concatenation as well; its
    # train-mode BatchNorm layers (DeepLab v3+) normalise every mini-batch with its own statistics and draw dropout masks per
    # mini-batch (engine.bn_train with tape.parts), i.e. the arithmetic of consecutive reference passes.
    def _graph_trunk(self, tape, x, in_h, in_w):
        """Run the trunk on `x` (Act, NHWC, ld 4); return [(feature Act, split_is_only_consumer), ...]."""
        raise NotImplementedError

    def _graph_head(self, tape, feats, in_h, in_w):
        """Run the head on the trunk features; return (low-res logits Act, align_corners)."""
        raise NotImplementedError

    def _trunk_module(self):
        raise NotImplementedError

    def _graph(self, tape, x, in_h, in_w):
        """Run the network on `x` (Act, NHWC, ld 4); return (low-res logits Act, align_corners)."""
        feats = self._graph_trunk(tape, x, in_h, in_w)
        return self._graph_head(tape, [f for f, _ in feats], in_h, in_w)

    def b2_trunk_is_batch_invariant(self):
        """True when no layer of the trunk mixes samples: every BatchNorm in eval mode (frozen), no active dropout."""
        for m in self._trunk_module().modules():
            name = type(m).__name__
            if m.training and ('BatchNorm' in name or ('Dropout' in name and getattr(m, 'p', 0) > 0)):
                return False
        return True

    # ---- explicit API used by the fused training step ---------------------------------------
    def b2_forward(self, x, record, input_grad=False):
        """x: (N,3,H,W) fp32 CUDA tensor.  Returns (logits (N,C,H,W) fp32, state for b2_backward).  `input_grad`: also record
        what d(loss)/d(x) needs (VAT, train_seg_semisup_vat_mt.py:247-268); b2_backward then returns that gradient."""
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError('expected an (N,3,H,W) image batch')
        K = get_kernels(self._n_split())
        if K.name == 'cuda' and not x.is_cuda:
            raise RuntimeError('B200 network received a CPU tensor: there is no CPU fallback (move the batch and the '
                               'network to a CUDA device)')
        x = x.detach().to(torch.float32)
        tape = E.Tape(K, enabled=record)
        E.prefold(K, self)                      # every frozen BatchNorm of the network in one launch
        xin = K.nchw_to_act(x, 4)
        xin.needs_grad = bool(input_grad and record)
        low, align = self._graph(tape, xin, x.shape[2], x.shape[3])
        logits = E.to_logits_nchw(tape, low, x.shape[2], x.shape[3], align)
        if not record:
            tape.discard()
            return logits, None
        return logits, _RunState(tape, low, align, xin if xin.needs_grad else None)

    def b2_forward_multi(self, xs, record):
        """Several mini-batches [(N_i,3,H,W)] of the same spatial size in one pass: the batch-invariant trunk runs once
        over their concatenation, the head once per mini-batch, in list order (so BatchNorm running statistics and
        dropout draws advance exactly as in consecutive `b2_forward` calls).  Returns ([logits_i], state)."""
        if not self.b2_trunk_is_batch_invariant():
            raise RuntimeError('b2_forward_multi needs a batch-invariant trunk (call freeze_batchnorm() first)')
        K = get_kernels(self._n_split())
        for x in xs:
            if x.dim() != 4 or x.shape[1] != 3 or x.shape[2:] != xs[0].shape[2:]:
                raise ValueError('expected (N,3,H,W) image batches of one spatial size')
            if K.name == 'cuda' and not x.is_cuda:
                raise RuntimeError('B200 network received a CPU tensor: there is no CPU fallback')
        sizes = [int(x.shape[0]) for x in xs]
        x_all = torch.cat([x.detach().to(torch.float32) for x in xs], dim=0)
        in_h, in_w = x_all.shape[2], x_all.shape[3]
        tape = E.Tape(K, enabled=record)
        E.prefold(K, self)
        xin = K.nchw_to_act(x_all, 4)
        xin.needs_grad = False
        feats = self._graph_trunk(tape, xin, in_h, in_w)
        logits, lows = [], []
        if self.b2_batched_head:
            # the head's convolutions also run once over all mini-batches (a 16-image head launch fills 3.46 of 4 waves of CTA
            # pairs, a 32-image one 6.92 of 7); its train-mode BatchNorms normalise per mini-batch (engine.bn_train, tape.parts)
            tape.parts = sizes
            low_all, align = self._graph_head(tape, [f for f, _ in feats], in_h, in_w)
            tape.parts = None
            for low in E.batch_split(tape, low_all, sizes):
                logits.append(E.to_logits_nchw(tape, low, in_h, in_w, align))
                lows.append((low, align))
        else:
            split = [E.batch_split(tape, f, sizes, delegate_gate=only) for f, only in feats]
            for i in range(len(xs)):
                low, align = self._graph_head(tape, [parts[i] for parts in split], in_h, in_w)
                logits.append(E.to_logits_nchw(tape, low, in_h, in_w, align))
                lows.append((low, align))
        if not record:
            tape.discard()
            return logits, None
        return logits, _MultiRunState(tape, lows)

    def b2_backward_multi(self, state, dlogits_list, scale_devs=None, cuts_for=None):
        """Backward of b2_forward_multi: one d(loss_i)/d(logits_i) (NCHW, optionally times a device scalar) per
        mini-batch; parameter gradients accumulate into .grad as the sum over the mini-batches.
        `cuts_for(ready)`: optional; given {id(parameter): node index at which its gradient is final} it returns the
        {node index: callable} hooks of Tape.backward (gradient-bucket all-reduces overlapped with the pass)."""
        if state is None or state.consumed:
            raise RuntimeError('this forward pass was not recorded or has already been back-propagated')
        state.consumed = True
        if scale_devs is None:
            scale_devs = [None] * len(dlogits_list)
        assert len(dlogits_list) == len(state.lows) == len(scale_devs)
        for (low, align), dl, sc in zip(state.lows, dlogits_list, scale_devs):
            E.seed_output_grad(state.tape, low, dl, align, scale_dev=sc)
        cuts = cuts_for(state.tape.param_ready_index()) if cuts_for is not None else None
        state.tape.backward(cuts)
        for low, _ in state.lows:
            low.grad = None
        state.tape, state.lows = None, None

    def b2_backward(self, state, dlogits, scale_dev=None, scale_host=1.0, param_grads=True):
        """Back-propagate d(loss)/d(logits) (NCHW, optionally to be multiplied by a device scalar) into
        the parameters' .grad (accumulating, like autograd).  `param_grads=False`: no parameter gradient is produced
        (torch.autograd.grad w.r.t. the input only).  Returns d(loss)/d(x) (N,3,H,W) if the pass was recorded with
        `input_grad=True`, else None."""
        if state is None or state.consumed:
            raise RuntimeError('this forward pass was not recorded or has already been back-propagated')
        state.consumed = True
        state.tape.param_grads = bool(param_grads)
        E.seed_output_grad(state.tape, state.low, dlogits, state.align, scale_dev=scale_dev, scale_host=scale_host)
        K = state.tape.K
        state.tape.backward()
        state.low.grad = None
        dx = None
        if state.xin is not None:
            if state.xin.grad is None:
                raise RuntimeError('the input gradient was requested but no gradient reached the image')
            dx = K.act_to_nchw(state.xin.grad)
            state.xin.grad = None
        state.tape, state.low, state.xin = None, None, None
        return dx

    # ---- nn.Module surface ------------------------------------------------------------------
    def forward(self, x, feature_maps=False, use_dropout=False):
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if not needs_grad:
            return self.b2_forward(x, record=False)[0]
        anchor = next(p for p in self.parameters() if p.requires_grad)
        return _B2Function.apply(x, anchor, self)

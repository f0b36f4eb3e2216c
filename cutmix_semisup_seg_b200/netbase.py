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
    __slots__ = ('tape', 'low', 'align', 'consumed')

    def __init__(self, tape, low, align):
        self.tape, self.low, self.align, self.consumed = tape, low, align, False


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

    # precision of the tensor-core convolutions: 'tf32' (one pass; what cuDNN does by default) or
    # '3xtf32' (operands split in hi/lo parts, three passes: ~fp32 accuracy, used by the parity tests)
    def __init__(self):
        super(B2SegNet, self).__init__()
        self.b2_precision = os.environ.get('B200SEG_PRECISION', 'tf32')

    def _n_split(self):
        return {'tf32': 1, '3xtf32': 3, '4xtf32': 4}[self.b2_precision]

    # ---- subclasses define the graph --------------------------------------------------------
    def _graph(self, tape, x, in_h, in_w):
        """Run the network on `x` (Act, NHWC, ld 4); return (low-res logits Act, align_corners)."""
        raise NotImplementedError

    # ---- explicit API used by the fused training step ---------------------------------------
    def b2_forward(self, x, record):
        """x: (N,3,H,W) fp32 CUDA tensor.  Returns (logits (N,C,H,W) fp32, state for b2_backward)."""
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError('expected an (N,3,H,W) image batch')
        K = get_kernels(self._n_split())
        if K.name == 'cuda' and not x.is_cuda:
            raise RuntimeError('B200 network received a CPU tensor: there is no CPU fallback (move the batch and the '
                               'network to a CUDA device)')
        x = x.detach().to(torch.float32)
        tape = E.Tape(K, enabled=record)
        xin = K.nchw_to_act(x, 4)
        xin.needs_grad = False
        low, align = self._graph(tape, xin, x.shape[2], x.shape[3])
        logits = E.to_logits_nchw(tape, low, x.shape[2], x.shape[3], align)
        if not record:
            tape.discard()
            return logits, None
        return logits, _RunState(tape, low, align)

    def b2_backward(self, state, dlogits, scale_dev=None, scale_host=1.0):
        """Back-propagate d(loss)/d(logits) (NCHW, optionally to be multiplied by a device scalar) into
        the parameters' .grad (accumulating, like autograd)."""
        if state is None or state.consumed:
            raise RuntimeError('this forward pass was not recorded or has already been back-propagated')
        state.consumed = True
        E.seed_output_grad(state.tape, state.low, dlogits, state.align, scale_dev=scale_dev, scale_host=scale_host)
        state.tape.backward()
        state.low.grad = None
        state.tape, state.low = None, None

    # ---- nn.Module surface ------------------------------------------------------------------
    def forward(self, x, feature_maps=False, use_dropout=False):
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if not needs_grad:
            return self.b2_forward(x, record=False)[0]
        anchor = next(p for p in self.parameters() if p.requires_grad)
        return _B2Function.apply(x, anchor, self)

"""NHWC activation views shared by the engine and the kernel backends."""
import torch


class Act(object):
    """A (N, H, W, C) fp32 activation stored channels-last inside `base` (shape (N, H, W, ld)):
    channels [off, off + c) of every pixel.  Slices of a wider buffer replace torch.cat."""

    __slots__ = ('base', 'n', 'h', 'w', 'c', 'ld', 'off', 'gate_on_grad', 'node', 'parent', 'pending',
                 'grad', 'grad_owned', 'name', 'needs_grad', 'fused_stats', 'grad_home')

    def __init__(self, base, n, h, w, c, ld=None, off=0, name=''):
        self.base = base
        self.n, self.h, self.w, self.c = n, h, w, c
        self.ld = c if ld is None else ld
        self.off = off
        self.gate_on_grad = False   # producer fused a ReLU: stored grads are already gated (g-form)
        self.node = None            # tape node that produced this activation
        self.parent = None          # wider buffer this is a slice of
        self.pending = 0            # consumers that still have to contribute to .grad
        self.grad = None            # Act holding d(loss)/d(this)
        self.grad_owned = False
        self.name = name
        self.needs_grad = True      # False for network inputs (no dgrad is computed into them)
        self.fused_stats = None     # column sums written by the epilogue that finished .grad (engine.Tape)
        self.grad_home = None       # pre-assigned gradient buffer (batch slices write into their parent's gradient)

    @staticmethod
    def alloc(n, h, w, c, device, ld=None, name=''):
        ld = (c + 3) // 4 * 4 if ld is None else ld      # 16-byte pixel pitch (TMA global-stride rule)
        return Act(torch.empty((n, h, w, ld), device=device, dtype=torch.float32), n, h, w, c, ld, 0, name)

    @property
    def ptr(self):
        return self.base.data_ptr() + 4 * self.off

    @property
    def rows(self):
        return self.n * self.h * self.w

    @property
    def device(self):
        return self.base.device

    def slice(self, off, c, name=''):
        s = Act(self.base, self.n, self.h, self.w, c, self.ld, self.off + off, name)
        s.parent = self
        return s

    def batch_slice(self, n0, n, name=''):
        """Samples [n0, n0 + n) as an independent root activation sharing this one's storage (N is the outermost
        dimension of NHWC, so the slice is contiguous)."""
        assert 0 <= n0 and n0 + n <= self.n
        return Act(self.base[n0:n0 + n], n, self.h, self.w, self.c, self.ld, self.off, name)

    def like(self, name=''):
        """Fresh dense buffer with the same logical shape."""
        return Act.alloc(self.n, self.h, self.w, self.c, self.device, name=name)

    def view4(self):
        """torch view (N, H, W, C) of the logical contents (strided when a slice)."""
        return self.base.view(self.n, self.h, self.w, self.ld)[..., self.off:self.off + self.c]

    def to_nchw(self):
        """Dense NCHW copy (plumbing for tests / feature-map export)."""
        return self.view4().permute(0, 3, 1, 2).contiguous()

    def __repr__(self):
        return 'Act({} n{} {}x{} c{} ld{} off{})'.format(self.name, self.n, self.h, self.w, self.c, self.ld, self.off)

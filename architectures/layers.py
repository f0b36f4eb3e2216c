"""Parameter-holder modules of the B200 networks.

The reference builds its networks from torch.nn layers whose forward passes are ATen/cuDNN calls.
Here the arithmetic lives in the engine (cutmix_semisup_seg_b200/engine.py -> libb200seg.so); these
modules only own the parameters/buffers under the SAME names, shapes and dtypes as the reference
`state_dict()` (so pretrained checkpoints, EMA key checks and saved models interoperate) plus the
layer hyper-parameters the engine needs.

Convolution weights keep the logical (Cout, Cin, kh, kw) shape but are stored channels-last, i.e.
physically (Cout, kh, kw, Cin) = the K-major "KRSC" operand layout the tensor-core GEMM reads through
TMA, so no per-step weight repacking is needed for the forward pass.
"""
import math

import torch
import torch.nn as nn


class B2Conv2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, bias=False):
        super(B2Conv2d, self).__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = (kernel_size, kernel_size)
        self.stride, self.padding, self.dilation = stride, padding, dilation
        w = torch.empty(out_channels, in_channels, kernel_size, kernel_size)
        self.weight = nn.Parameter(w.contiguous(memory_format=torch.channels_last))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        # torch.nn.Conv2d default: kaiming_uniform(a=sqrt(5)) weights, U(-1/sqrt(fan_in), ..) bias
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in = self.in_channels * self.kernel_size[0] * self.kernel_size[1]
            bound = 1.0 / math.sqrt(fan_in)
            nn.init.uniform_(self.bias, -bound, bound)

    def extra_repr(self):
        return '{}, {}, kernel_size={}, stride={}, padding={}, dilation={}, bias={}'.format(
            self.in_channels, self.out_channels, self.kernel_size, self.stride, self.padding, self.dilation,
            self.bias is not None)

    def forward(self, x):
        raise RuntimeError('B2Conv2d is a parameter holder; call the owning network (engine executes the graph)')


class B2BatchNorm2d(nn.Module):
    """BatchNorm2d state: weight, bias, running_mean, running_var, num_batches_tracked (int64)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1):
        super(B2BatchNorm2d, self).__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))
        self.register_buffer('running_mean', torch.zeros(num_features))
        self.register_buffer('running_var', torch.ones(num_features))
        self.register_buffer('num_batches_tracked', torch.tensor(0, dtype=torch.long))

    def extra_repr(self):
        return '{}, eps={}, momentum={}'.format(self.num_features, self.eps, self.momentum)

    def forward(self, x):
        raise RuntimeError('B2BatchNorm2d is a parameter holder; call the owning network')


class B2Dropout(nn.Module):
    """Dropout(p) marker.  Masks come from the counter-based generator in libb200seg.so, or from an
    explicit queue (`inject`) when a test needs the same masks in two implementations."""

    def __init__(self, p=0.5):
        super(B2Dropout, self).__init__()
        self.p = p
        self._injected = []
        self._dev_counter = None
        self.seed = 0x5EED

    def inject(self, masks):
        """Queue explicit keep-masks (NHWC float 0/1 tensors), consumed in order by the next forwards."""
        self._injected = list(masks)

    def next_mask(self, kernels, n, h, w, c, device):
        if self._injected:
            m = self._injected.pop(0).to(device=device, dtype=torch.float32).contiguous()
            assert tuple(m.shape) == (n, h, w, c), 'injected dropout mask has shape {}'.format(tuple(m.shape))
            return m
        # the stream position lives on the device so that a CUDA-graph replay still draws a fresh mask
        if self._dev_counter is None or self._dev_counter.device != torch.device(device):
            self._dev_counter = torch.zeros((1,), dtype=torch.int64, device=device)
        self._dev_counter += n * h * w * c
        return kernels.dropout_mask(n, h, w, c, self.p, self.seed, 0, device, offset_dev=self._dev_counter)

    def __getstate__(self):
        d = dict(self.__dict__)
        d['_injected'] = []
        d['_dev_counter'] = None
        return d


class B2Marker(nn.Module):
    """Parameter-less placeholder that keeps Sequential indices identical to the reference
    (ReLU / pooling layers have no state)."""

    def __init__(self, kind=''):
        super(B2Marker, self).__init__()
        self.kind = kind

    def extra_repr(self):
        return self.kind

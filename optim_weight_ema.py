"""Drop-in for the reference's `optim_weight_ema` module (optim_weight_ema.py:4-25).

`EMAWeightOptimizer(target_net, source_net, ema_alpha)` keeps the reference's attributes
(`target_params`, `source_params`: every float32 `state_dict()` tensor, i.e. parameters AND BatchNorm
running statistics; int64 `num_batches_tracked` is skipped) and semantics (constructor copies
source -> target and checks the key sets; `step()` does `t = t*alpha + s*(1-alpha)`), but `step()`
is ONE fused multi-tensor CUDA kernel (`b2_ema_step` over a device table of chunks) instead of 3 launches x ~567 tensors.  The arithmetic is
bit-exact with the reference: three separate fp32 roundings, no FMA.
"""
import numpy as np
import torch

_CHUNK = 16384  # elements per thread block, must match EMA_CHUNK in csrc/elementwise.cu


class EMAWeightOptimizer(object):
    def __init__(self, target_net, source_net, ema_alpha, host_arithmetic=False):
        """`host_arithmetic=True` (new, opt-in): both networks live on the CPU and `step()` runs in host arithmetic -- only the
        toy-2D plumbing configuration asks for it (`toy2d_train.py --device cpu`, BASELINE config 1).  Without it CPU tensors are
        refused: the segmentation hot path has no CPU fallback."""
        self.host_arithmetic = bool(host_arithmetic)
        self.target_net = target_net
        self.source_net = source_net
        self.ema_alpha = ema_alpha
        self.target_params = [p for p in target_net.state_dict().values() if p.dtype == torch.float]
        self.source_params = [p for p in source_net.state_dict().values() if p.dtype == torch.float]

        for tgt_p, src_p in zip(self.target_params, self.source_params):
            tgt_p[...] = src_p[...]

        target_keys = set(target_net.state_dict().keys())
        source_keys = set(source_net.state_dict().keys())
        if target_keys != source_keys:
            raise ValueError('Source and target networks do not have the same state dict keys; do they have '
                             'different architectures?')
        self._table = None
        self._table_key = None
        self._n_chunks = 0

    # ------------------------------------------------------------------------------------------
    def _build_table(self, backend_device):
        ptr_key = tuple(p.data_ptr() for p in self.target_params) + tuple(p.data_ptr() for p in self.source_params)
        if self._table is not None and ptr_key == self._table_key:
            return
        rows = []
        for t, s in zip(self.target_params, self.source_params):
            if t.numel() != s.numel():
                raise ValueError('EMA: tensor size mismatch between target and source')
            if not (_dense(t) and _dense(s) and t.stride() == s.stride()):
                raise ValueError('EMA: tensors must be dense with identical strides')
            n = t.numel()
            tp, sp = t.data_ptr(), s.data_ptr()
            for off in range(0, n, _CHUNK):
                rows.append((tp + 4 * off, sp + 4 * off, min(_CHUNK, n - off)))
        arr = np.array(rows, dtype=np.int64).reshape(-1, 3)   # struct b2_ema_chunk = {ptr, ptr, int64}
        self._table = torch.from_numpy(arr).to(backend_device)
        self._table_key = ptr_key
        self._n_chunks = arr.shape[0]

    def _host_step(self):
        """Both networks live on the CPU (`toy2d_train.py --device cpu`, BASELINE config 1: the reference's CPU plumbing
        configuration): the update in host arithmetic, with the reference's three fp32 roundings (optim_weight_ema.py:21-25).
        Opt-in through the constructor (a device TARGET the caller chose for the whole model, not a fallback): CUDA tensors
        never take this path."""
        if not all(p.device.type == 'cpu' for p in self.target_params + self.source_params):
            raise RuntimeError('EMAWeightOptimizer(host_arithmetic=True) needs every tensor of both networks on the CPU')
        one_minus_alpha = 1.0 - self.ema_alpha
        for tgt_p, src_p in zip(self.target_params, self.source_params):
            tgt_p.mul_(self.ema_alpha)
            tgt_p.add_(src_p * one_minus_alpha)

    def step(self):
        if len(self.target_params) == 0:
            return
        dev = self.target_params[0].device
        if self.host_arithmetic:
            return self._host_step()
        if dev.type != 'cuda' or any(p.device != dev for p in self.target_params + self.source_params):
            raise RuntimeError('EMAWeightOptimizer (B200 hot path) needs all tensors on one CUDA device; '
                               'there is no CPU fallback')
        from cutmix_semisup_seg_b200 import ops, engine
        be = ops.default_backend()
        engine.invalidate_caches()      # the kernel writes the teacher through raw pointers (no torch version bump)
        with torch.cuda.device(dev):
            self._build_table(dev)
            be.ema_step_table(self._table, self._n_chunks, self.ema_alpha)


def _dense(t):
    """True if the tensor's elements occupy one gap-free block (any dim order)."""
    return t.is_contiguous() or t.is_contiguous(memory_format=torch.channels_last) or \
        t.numel() == t.untyped_storage().nbytes() // 4 or _is_permuted_dense(t)


def _is_permuted_dense(t):
    sizes_strides = sorted(((st, sz) for sz, st in zip(t.shape, t.stride()) if sz > 1))
    expect = 1
    for st, sz in sizes_strides:
        if st != expect:
            return False
        expect *= sz
    return True

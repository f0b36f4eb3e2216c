"""Fused student-optimiser + teacher-EMA step (reference train_seg_semisup_mask_mt.py:90-100, 465-467).

`FusedOptimizer` is a `torch.optim.Optimizer` (so the reference's LR schedulers, `zero_grad()`, `param_groups`,
`state_dict()` keep working) whose `step()` is ONE launch of `b2_opt_ema_step` (csrc/optim.cu) over every parameter of
every group -- and, when an `EMAWeightOptimizer` is passed, over the teacher's EMA copies in the same pass.

The reference's parameter groups are taken literally: DeepLab v2's "pretrained" group repeats tensors (314 entries,
104 unique, deeplab2.py:224-230) and torch.optim's per-tensor loop -- what the reference ran -- therefore applies k
sequential updates per step to a tensor listed k times, with a shared state and its step counter advancing by k.  The
kernel reproduces exactly that (torch's own multi-tensor kernels do not), in fp32 with the operation order of
`torch.optim.adam._single_tensor_adam` / `sgd._single_tensor_sgd`.
"""
import warnings

import numpy as np
import torch

from . import lib as L

_CHUNK = 8192           # B2_OPT_CHUNK in include/b200seg.h
_MAX_K = 8              # B2_OPT_MAX_K
_LR_SLOTS = 8           # pinned staging slots of the learning-rate upload (see FusedOptimizer.upload_lr)


class FusedOptimizer(torch.optim.Optimizer):
    def __init__(self, params, opt_type='adam', lr=1e-3, betas=(0.9, 0.999), eps=1e-8, momentum=0.0, nesterov=False,
                 weight_decay=0.0):
        if opt_type not in ('adam', 'sgd'):
            raise ValueError('Unknown opt_type {}'.format(opt_type))
        if opt_type == 'adam' and weight_decay != 0.0:
            raise ValueError('the fused Adam has no weight decay (the reference uses none)')
        if nesterov and momentum <= 0.0:
            raise ValueError('Nesterov momentum requires a momentum')
        defaults = dict(lr=lr, betas=betas, eps=eps, momentum=momentum, nesterov=nesterov, weight_decay=weight_decay)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')          # duplicated parameters are the reference's behaviour, kept on purpose
            super(FusedOptimizer, self).__init__(params, defaults)
        self.opt_type = opt_type
        g0 = self.param_groups[0]
        for g in self.param_groups[1:]:
            for key in ('betas', 'eps', 'momentum', 'nesterov', 'weight_decay'):
                if g[key] != g0[key]:
                    raise ValueError('the fused optimiser needs the same {} in every group (only lr may differ)'.format(key))
        # unique parameters in first-appearance order, their multiplicity and (first) group
        self._uniq, self._mult, self._group = [], {}, {}
        for gi, g in enumerate(self.param_groups):
            for p in g['params']:
                if not p.requires_grad:
                    continue
                if id(p) not in self._mult:
                    self._uniq.append(p)
                    self._mult[id(p)] = 0
                    self._group[id(p)] = gi
                elif self._group[id(p)] != gi:
                    raise ValueError('a parameter appears in more than one group')
                self._mult[id(p)] += 1
        if any(k > _MAX_K for k in self._mult.values()):
            raise ValueError('a parameter is listed more than {} times'.format(_MAX_K))
        if not self._uniq:
            raise ValueError('optimizer got an empty parameter list')
        dev = self._uniq[0].device
        if dev.type != 'cuda' or any(p.device != dev or p.dtype != torch.float32 for p in self._uniq):
            raise RuntimeError('FusedOptimizer (B200 hot path) needs fp32 parameters on one CUDA device; there is no '
                               'CPU fallback')
        self._dev = dev
        total = sum(p.numel() for p in self._uniq)
        self._m = torch.zeros(total, device=dev, dtype=torch.float32)
        self._v = torch.zeros(total, device=dev, dtype=torch.float32) if opt_type == 'adam' else None
        self._iter = torch.zeros((), device=dev, dtype=torch.int64)
        off = 0
        self._off = {}
        for i, p in enumerate(self._uniq):
            n = p.numel()
            st = self.state[p]
            st['exp_avg' if opt_type == 'adam' else 'momentum_buffer'] = self._m[off:off + n]
            if opt_type == 'adam':
                st['exp_avg_sq'] = self._v[off:off + n]
            if i == 0:
                st['b2_steps_done'] = self._iter          # shared step counter (every tensor steps every iteration)
            self._off[id(p)] = off
            off += n
        # Learning rates reach the kernel through a RING of pinned host slots: the host-to-device copy is asynchronous, so a
        # single staging buffer could be overwritten with iteration i+1's rate (poly / cosine schedules change it every
        # iteration) before iteration i's copy has executed.  Each slot carries an event recorded after its copy; a slot is
        # rewritten only once that event has completed (the host blocks only if the GPU is _LR_SLOTS iterations behind).
        self._lr_ring = [torch.zeros(len(self.param_groups), dtype=torch.float64).pin_memory() for _ in range(_LR_SLOTS)]
        self._lr_events = [None] * _LR_SLOTS
        self._lr_next = 0
        self._lr_dev = torch.zeros(len(self.param_groups), device=dev, dtype=torch.float64)
        self._table = None
        self._table_key = None
        self._n_chunks = 0
        dup = any(k > 1 for k in self._mult.values())
        self.b2_fused = True
        self.b2_note = '{} (fused multi-tensor sm_100a kernel{})'.format(
            opt_type, ', k sequential updates for the duplicated reference group' if dup else '')

    # ------------------------------------------------------------------------------------------
    def upload_lr(self):
        """Copy the groups' current learning rates to the device buffer the kernel reads, ordered on the current stream
        before the next step launch / CUDA-graph replay (the captured step contains no copy of its own: call this before
        every replay)."""
        j = self._lr_next
        self._lr_next = (j + 1) % _LR_SLOTS
        ev = self._lr_events[j]
        if ev is not None:
            ev.synchronize()                   # the copy that last read this slot has executed
        lr = self._lr_ring[j].numpy()
        for i, g in enumerate(self.param_groups):
            lr[i] = float(g['lr'])
        with torch.cuda.device(self._dev):
            self._lr_dev.copy_(self._lr_ring[j], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        self._lr_events[j] = ev

    @staticmethod
    def _dense(t):
        n, expect = t.numel(), 1
        for st, sz in sorted((st, sz) for sz, st in zip(t.shape, t.stride()) if sz > 1):
            if st != expect:
                return False
            expect *= sz
        return n > 0

    def _build_table(self, ema):
        params = [p for p in self._uniq if p.grad is not None]
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in params)
        pairs = []
        if ema is not None:
            pairs = list(zip(ema.target_params, ema.source_params))
            key += tuple((t.data_ptr(), s.data_ptr()) for t, s in pairs)
        if self._table is not None and key == self._table_key:
            return
        teacher_of = {}
        for t, s in pairs:
            if t.numel() != s.numel() or not (self._dense(t) and self._dense(s) and t.stride() == s.stride()):
                raise ValueError('EMA: tensors must be dense with identical sizes and strides')
            if t.device != self._dev or s.device != self._dev:
                raise RuntimeError('EMA tensors must live on the optimiser\'s CUDA device')
            teacher_of[s.data_ptr()] = t
        rows = []

        def add(p_ptr, g_ptr, m_ptr, v_ptr, t_ptr, n, group, k):
            for o in range(0, n, _CHUNK):
                c = min(_CHUNK, n - o)
                b = 4 * o
                rows.append((p_ptr + b, g_ptr + b if g_ptr else 0, m_ptr + b if m_ptr else 0, v_ptr + b if v_ptr else 0,
                             t_ptr + b if t_ptr else 0, c + (group << 32) + (k << 48)))
        stepped = set()
        for p in params:
            if not self._dense(p) or p.grad.stride() != p.stride() or not p.grad.is_cuda:
                raise ValueError('parameters must be dense and gradients laid out like them')
            off = self._off[id(p)]
            t = teacher_of.get(p.data_ptr())
            if t is not None and t.stride() != p.stride():
                raise ValueError('EMA target laid out differently from its source')
            add(p.data_ptr(), p.grad.data_ptr(), self._m.data_ptr() + 4 * off,
                self._v.data_ptr() + 4 * off if self._v is not None else 0, t.data_ptr() if t is not None else 0,
                p.numel(), self._group[id(p)], self._mult[id(p)])
            stepped.add(p.data_ptr())
        for t, s in pairs:                     # EMA-only state: BatchNorm buffers, frozen / grad-less parameters
            if s.data_ptr() not in stepped:
                add(s.data_ptr(), 0, 0, 0, t.data_ptr(), s.numel(), 0, 0)
        arr = np.array(rows, dtype=np.int64).reshape(-1, 6)        # struct b2_opt_chunk (48 bytes)
        self._table = torch.from_numpy(arr).to(self._dev)
        self._table_key = key
        self._n_chunks = arr.shape[0]

    @torch.no_grad()
    def step(self, closure=None, ema=None):
        """One optimiser step; with `ema` (an optim_weight_ema.EMAWeightOptimizer whose source network owns these
        parameters) the teacher's EMA step runs in the same launch -- do not call `ema.step()` as well."""
        if closure is not None:
            raise RuntimeError('closures are not supported by the fused optimiser')
        from . import engine, ops
        self._build_table(ema)
        if not torch.cuda.is_current_stream_capturing():
            self.upload_lr()            # (a captured step reads what upload_lr() delivered before the replay)
        g = self.param_groups[0]
        with torch.cuda.device(self._dev):
            alpha = float(ema.ema_alpha) if ema is not None else 0.0
            ops.default_backend()._call('b2_opt_ema_step', self._table.data_ptr(), self._n_chunks, self._lr_dev.data_ptr(), self._iter.data_ptr(),
                   0 if self.opt_type == 'adam' else 1, float(g['betas'][0]), float(g['betas'][1]), float(g['eps']),
                   float(g['momentum']), float(g['weight_decay']), int(bool(g['nesterov'])), int(ema is not None),
                   alpha, 1.0 - alpha, L.stream_ptr())
        engine.invalidate_caches()      # parameters were written through raw pointers (no torch version bump)
        ops.default_backend().launches += 1       # (the call above counted one; the step-counter tick is the second launch)
        return None

    def load_state_dict(self, state_dict):
        """Standard torch format; the loaded moments are copied into this optimiser's flat buffers (the kernel's chunk
        table points into them)."""
        super(FusedOptimizer, self).load_state_dict(state_dict)
        mkey = 'exp_avg' if self.opt_type == 'adam' else 'momentum_buffer'
        with torch.no_grad():
            for i, p in enumerate(self._uniq):
                st, off, n = self.state[p], self._off[id(p)], p.numel()
                for key, flat in ((mkey, self._m), ('exp_avg_sq', self._v)):
                    if flat is None:
                        continue
                    view = flat[off:off + n]
                    if key in st and torch.is_tensor(st[key]) and st[key].data_ptr() != view.data_ptr():
                        view.copy_(st[key].reshape(-1))
                    st[key] = view
                if i == 0:
                    if 'b2_steps_done' in st and st['b2_steps_done'].data_ptr() != self._iter.data_ptr():
                        self._iter.copy_(st['b2_steps_done'])
                    st['b2_steps_done'] = self._iter

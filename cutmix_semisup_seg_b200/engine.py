"""Define-by-run executor for the DeepLab forward/backward passes on the B200 kernels.

The reference relies on torch.autograd over ~2000 ATen launches per iteration
(train_seg_semisup_mask_mt.py:299-301, 355-358, 459).  Here each network layer is one fused kernel
(conv + folded frozen-BN + residual + ReLU in the GEMM epilogue) and the backward pass is an
explicit tape: every node knows its own gradient kernels, and ReLU / residual / BN-scale backward
steps are folded into the producing dgrad epilogue, the transposed-weight pass or the wgrad
epilogue instead of being separate HBM passes.

Gradient bookkeeping
  * `Act.pending` counts consumers that still owe a contribution to `Act.grad`.
  * A convolution input-gradient that is the LAST contribution to an activation adds the partial
    sum (`addend`) and applies the producer's ReLU gate in its epilogue (g-form gradient).
  * Residual branches contribute by aliasing (no copy, no add kernel).
  * The same epilogue that finishes the gradient of a frozen-BN layer's output also writes the column sums the
    BN affine parameters need (sum g, sum g*(y - residual)); the separate reduction pass over g and y is skipped.
"""
import os
import weakref

import torch

from .acts import Act

# ---------------------------------------------------------------------------------------- derived-weight caches
# Folded frozen-BN constants and transposed (dgrad) weights depend only on the parameters, so the two student passes
# and the two teacher passes of one iteration share them.  An entry is valid while (generation, tensor versions)
# are unchanged: torch bumps `_version` on every in-place update (optimiser step, load_state_dict), and code that
# writes parameters through raw pointers (the fused EMA kernel) or wants a fresh start (every training step, so that
# a CUDA-graph capture always contains the producing kernels) calls `invalidate_caches()`.
_GENERATION = [0]
_FOLD_CACHE = weakref.WeakKeyDictionary()
_WT_CACHE = weakref.WeakKeyDictionary()


def invalidate_caches():
    _GENERATION[0] += 1


# ---------------------------------------------------------------------------------------- multi-tensor prefills
# The per-layer kernels behind the two caches above are ~1 us of work each (208 BatchNorm folds and 113 weight transposes per
# cfg3 iteration: ~2.3 ms of launch-bound graph nodes).  `prefold` / `pretranspose` compute all entries a pass will ask for in
# ONE launch each (csrc/multi.cu, bit-identical arithmetic) and park them in the caches; the per-layer code paths stay as the
# fallback for anything the prefill did not cover.  Outputs live in persistent per-network buffers and the device tables are
# rebuilt only when a pointer changes, so a CUDA-graph capture sees the same addresses as its warm-up pass.
_PREFILL = weakref.WeakKeyDictionary()        # network -> {'fold': (...), 'wt': {...}}
_BN_ENTRY = 56
_TR_ENTRY = 48


def _device_table(records, dtype_fields, device):
    import numpy as np
    arr = np.zeros(len(records), dtype=np.dtype(dtype_fields, align=False))
    for i, rec in enumerate(records):
        arr[i] = rec
    return torch.from_numpy(arr.view(np.uint8).copy()).to(device)


def prefold(K, net):
    """Fold every eval-mode BatchNorm of `net` (scale = gamma / sqrt(var + eps), shift = beta - mean * scale) in one launch."""
    if not getattr(K, 'multi_tensor', False):
        return
    bns = [m for m in net.modules() if type(m).__name__ == 'B2BatchNorm2d' and not m.training]
    if not bns:
        return
    hit = _FOLD_CACHE.get(bns[0])
    key0 = (_GENERATION[0], K.name, _versions(bns[0].weight, bns[0].bias, bns[0].running_mean, bns[0].running_var))
    if hit is not None and hit[0] == key0:
        return                                   # this generation has been folded already (second pass of the iteration)
    st = _PREFILL.setdefault(net, {})
    ptrs = tuple((m.weight.data_ptr(), m.bias.data_ptr(), m.running_mean.data_ptr(), m.running_var.data_ptr()) for m in bns)
    fold = st.get('fold')
    if fold is None or fold[0] != ptrs:
        dev = bns[0].weight.device
        total = sum(m.num_features for m in bns)
        buf = torch.empty((2, total), device=dev, dtype=torch.float32)
        recs, views, off = [], [], 0
        for m in bns:
            c = m.num_features
            sc, sh = buf[0, off:off + c], buf[1, off:off + c]
            recs.append((m.weight.data_ptr(), m.bias.data_ptr(), m.running_mean.data_ptr(), m.running_var.data_ptr(),
                         sc.data_ptr(), sh.data_ptr(), c, float(m.eps)))
            views.append((sc, sh))
            off += c
        fields = [('gamma', 'u8'), ('beta', 'u8'), ('mean', 'u8'), ('var', 'u8'), ('scale', 'u8'), ('shift', 'u8'), ('c', 'i4'),
                  ('eps', 'f4')]
        table = _device_table(recs, fields, dev)
        assert table.numel() == _BN_ENTRY * len(recs)
        fold = (ptrs, buf, views, table, max(m.num_features for m in bns))
        st['fold'] = fold
    _, buf, views, table, max_c = fold
    K.bn_fold_multi(table, len(bns), max_c)
    for m, (sc, sh) in zip(bns, views):
        key = (_GENERATION[0], K.name, _versions(m.weight, m.bias, m.running_mean, m.running_var))
        _FOLD_CACHE[m] = (key, sc, sh)


def pretranspose(K, nodes):
    """Produce the (cin, taps, pad4(cout)) dgrad operands of every convolution node of a recorded pass in one launch."""
    if not getattr(K, 'multi_tensor', False) or getattr(K, 'n_split', 1) != 1:
        return
    todo = []
    for node in nodes:
        if not isinstance(node, ConvNode) or node.col_src is not None or node.x is None or not node.x.needs_grad:
            continue
        conv, scale = node.conv, node.scale
        w = conv.weight
        key = (_GENERATION[0], K.name, 1, _versions(w), None if scale is None else _versions(scale))
        hit = _WT_CACHE.get(conv)
        if hit is not None and hit[0] == key:
            continue
        todo.append((conv, scale, key, node.geom))
    if len(todo) < 2:
        return
    seen, uniq = set(), []
    for item in todo:
        if id(item[0]) not in seen:
            seen.add(id(item[0]))
            uniq.append(item)
    owner = uniq[0][0]
    st = _PREFILL.setdefault(owner, {})
    ptrs = tuple((c.weight.data_ptr(), 0 if s is None else s.data_ptr()) for c, s, _, _ in uniq)
    wt = st.get('wt')
    if wt is None or wt[0] != ptrs:
        dev = owner.weight.device
        recs, outs, blocks = [], [], 0
        for conv, scale, _, geom in uniq:
            cout, kh, kw, cin = geom[0], geom[1], geom[2], geom[3]
            t, ldd = kh * kw, (cout + 3) // 4 * 4
            out = torch.empty((cin, t, ldd), device=dev, dtype=torch.float32)
            recs.append((conv.weight.data_ptr(), out.data_ptr(), 0 if scale is None else scale.data_ptr(), cout, t, cin, ldd, blocks))
            blocks += ((cin + 31) // 32) * ((ldd + 31) // 32) * t
            outs.append((out, ldd))
        fields = [('src', 'u8'), ('dst', 'u8'), ('scale', 'u8'), ('a', 'i4'), ('t', 'i4'), ('b', 'i4'), ('ldd', 'i4'),
                  ('block_begin', 'i8')]
        table = _device_table(recs, fields, dev)
        assert table.numel() == _TR_ENTRY * len(recs)
        wt = (ptrs, outs, table, blocks)
        st['wt'] = wt
    _, outs, table, blocks = wt
    K.transpose_w_multi(table, len(uniq), blocks)
    for (conv, scale, key, _), (out, ldd) in zip(uniq, outs):
        _WT_CACHE[conv] = (key, out, ldd, scale)


def _versions(*tensors):
    return tuple((t.data_ptr(), t._version) for t in tensors)


class Tape(object):
    # `parts`: sizes of the mini-batches concatenated along the batch dimension while the graph of a multi-batch pass is recorded
    # (netbase.b2_forward_multi); train-mode BatchNorm then normalises every mini-batch with its OWN statistics (bn_train), which
    # is all that distinguishes one pass over [x1 ; x2] from consecutive passes over x1 and x2
    parts = None

    def __init__(self, kernels, enabled):
        self.K = kernels
        self.enabled = enabled
        self.nodes = []
        # False: back-propagate to the network INPUT only (VAT's perturbation direction, torch.autograd.grad(loss, eps) in
        # train_seg_semisup_vat_mt.py:265-268): no parameter receives a gradient, whatever its requires_grad says
        self.param_grads = True

    def wants(self, p):
        """Does this backward pass produce a gradient for parameter p?"""
        return self.param_grads and p.requires_grad

    def record(self, node, inputs):
        if not self.enabled:
            y = getattr(node, 'y', None)
            if y is not None:
                y.node = None
            for k in list(vars(node).keys()):
                setattr(node, k, None)
            return
        for a in inputs:
            root = a
            while root.parent is not None:
                root = root.parent
            root.pending += 1
        self.nodes.append(node)

    # ---------------------------------------------------------------- gradient accumulation
    @staticmethod
    def _new_grad(t):
        """Fresh gradient buffer for t: its pre-assigned home (batch slices, see batch_split) or a new allocation."""
        return t.grad_home if t.grad_home is not None else t.like()

    def _finish(self, t, fused_gate):
        if t.pending == 0 and t.grad_home is not None and t.grad is not None and t.grad is not t.grad_home:
            self.K.copy_act(t.grad_home, t.grad)         # (aliased residual gradient: move it home)
            t.grad, t.grad_owned = t.grad_home, True
        if t.pending == 0 and t.gate_on_grad and not fused_gate and t.grad is not None:
            if not t.grad_owned:
                owned = self._new_grad(t)
                self.K.copy_act(owned, t.grad)
                t.grad, t.grad_owned = owned, True
            self.K.relu_gate(t.grad, t)

    def _stats_request(self, t):
        """Residual Act to subtract (or False) if the epilogue finishing d/d(t) should also emit the frozen-BN
        parameter-gradient statistics of the node that produced t; None = no statistics wanted."""
        node = t.node
        if not isinstance(node, ConvNode) or node.bn is None or not self.wants(node.bn.weight):
            return None
        if self.wants(node.conv.weight):
            # dgamma comes from <W, dW> (ConvNode.backward): only sum_pix g is wanted, no pass over y / the residual
            return False if self.K.stats_ok(t) else None
        if not self.K.stats_ok(t) or (node.residual is not None and not self.K.stats_ok(node.residual)):
            return None
        return node.residual if node.residual is not None else False

    def contribute_kernel(self, t, launch, fusable):
        """launch(dst, accumulate, addend, gate, stats_sub) writes one consumer's contribution to d/d(t); stats_sub is
        None (no statistics), False (statistics, nothing subtracted) or the residual Act; it returns the statistics."""
        assert t.parent is None, 'gradients are accumulated on root activations'
        t.pending -= 1
        last = t.pending == 0
        fused_gate = False
        if fusable:
            gate = t if (last and t.gate_on_grad) else None
            fused_gate = gate is not None
            want = self._stats_request(t) if fused_gate else None
            if t.grad is None:
                dst = self._new_grad(t)
                st = launch(dst, False, None, gate, want)
            elif t.grad_owned and not last:
                dst = t.grad
                st = launch(dst, True, None, None, None)
            elif t.grad_owned:
                dst = t.grad
                st = launch(dst, False, dst, gate, want)   # epilogue reads the partial before overwriting it
            else:
                dst = self._new_grad(t)
                st = launch(dst, False, t.grad, gate, want)
            t.grad, t.grad_owned = dst, True
            t.fused_stats = st if want is not None else None
        else:
            if t.grad is None:
                dst = self._new_grad(t)
                launch(dst, False, None, None, None)
                t.grad, t.grad_owned = dst, True
            else:
                if not t.grad_owned:
                    owned = self._new_grad(t)
                    self.K.copy_act(owned, t.grad)
                    t.grad, t.grad_owned = owned, True
                launch(t.grad, True, None, None, None)
        self._finish(t, fused_gate)

    def contribute_tensor(self, t, g):
        """d/d(t) += g where g is an existing read-only gradient buffer (residual pass-through)."""
        assert t.parent is None
        t.pending -= 1
        if t.grad is None:
            t.grad, t.grad_owned = g, False
        elif t.grad_owned:
            self.K.copy_act(t.grad, g, accumulate=True)
        else:
            owned = self._new_grad(t)
            self.K.copy_act(owned, t.grad)
            self.K.copy_act(owned, g, accumulate=True)
            t.grad, t.grad_owned = owned, True
        self._finish(t, False)

    def contribute_slice(self, t, launch):
        """d/d(t) += one consumer's contribution where t is a channel SLICE of a wider root activation (DenseNet: every
        dense layer reads a prefix of the block's concatenation).  launch(dst, accumulate) adds into dst, a view of the
        root's gradient restricted to t's channels; the root's gradient is created zero-filled on first use."""
        if t.parent is None:
            return self.contribute_kernel(t, lambda dst, accumulate, addend, gate, stats: launch(dst, accumulate), False)
        root, off = t, 0
        while root.parent is not None:
            off += root.off - root.parent.off
            root = root.parent
        root.pending -= 1
        if root.grad is None:
            root.grad, root.grad_owned = self._new_grad(root), True
            self.K.fill_act(root.grad, 0.0)
        elif not root.grad_owned:
            owned = self._new_grad(root)
            self.K.copy_act(owned, root.grad)
            root.grad, root.grad_owned = owned, True
        launch(root.grad.slice(off, t.c), True)
        self._finish(root, False)

    def skip(self, t):
        root = t
        while root.parent is not None:
            root = root.parent
        root.pending -= 1
        self._finish(root, False)

    def param_ready_index(self):
        """{id(parameter): index of the EARLIEST recorded node that contributes to its gradient}.  The backward pass runs
        the nodes from the last to the first, so a parameter's gradient is final once the node with that index has run
        (parameters of layers applied to several mini-batches of one pass are accumulated by several nodes)."""
        ready = {}
        for i, node in enumerate(self.nodes):
            mods = [getattr(node, 'conv', None), getattr(node, 'bn', None)]
            for m in mods:
                if m is None:
                    continue
                for p in (getattr(m, 'weight', None), getattr(m, 'bias', None)):
                    if p is not None and id(p) not in ready:
                        ready[id(p)] = i
        return ready

    def backward(self, cuts=None):
        """Run the recorded nodes in reverse.  Each node is released as soon as it has run: its output's
        gradient and its saved activations are dropped (Act <-> node reference cycles are broken explicitly so
        tens of GB of activations are freed by reference counting, not by the cyclic GC).
        `cuts`: {node index: callable} -- called right after the node with that index has run (the training step launches
        the all-reduce of a gradient bucket there, see step.MeanTeacherStep)."""
        nodes, self.nodes = self.nodes, []
        pretranspose(self.K, nodes)              # all dgrad operands of this pass in one launch
        while nodes:
            node = nodes.pop()
            index = len(nodes)
            node.backward(self)
            if cuts and index in cuts:
                cuts[index]()
            y = getattr(node, 'y', None)
            if y is not None:
                y.node = None
                if y.parent is None:
                    y.grad = None
            for k in list(vars(node).keys()):
                setattr(node, k, None)
        if cuts and -1 in cuts:
            cuts[-1]()                           # "after the whole pass"

    def discard(self):
        for node in self.nodes:
            y = getattr(node, 'y', None)
            if y is not None:
                y.node = None
            for k in list(vars(node).keys()):
                setattr(node, k, None)
        self.nodes = []


def out_grad(act):
    """Gradient buffer of an activation (a view into the parent's gradient for slices)."""
    if act.parent is None:
        return act.grad
    root, off = act, 0
    while root.parent is not None:
        off += root.off - root.parent.off
        root = root.parent
    if root.grad is None:
        return None
    g = root.grad.slice(off, act.c)
    return g


def param_grad(p):
    """(grad tensor laid out like p, accumulate?)  Creates p.grad on first use.  The gradient kernels write through raw
    pointers in the PARAMETER's storage order, so an existing .grad with other strides (a user-assigned NCHW-contiguous
    tensor) is refused instead of being filled in the wrong order."""
    if p.grad is None:
        p.grad = torch.empty_like(p)
        return p.grad, False
    if p.grad.stride() != p.stride() or p.grad.dtype != torch.float32:
        raise RuntimeError('parameter gradient of shape {} has strides {} / dtype {}, the parameter has strides {}: the B200 '
                           'kernels need .grad laid out like the parameter (set it to None and let the backward pass create it)'
                           .format(tuple(p.shape), p.grad.stride(), p.grad.dtype, p.stride()))
    return p.grad, True


def check_weight_layout(w):
    """Convolution weights are logical (Cout, Cin, kh, kw) tensors STORED channels-last = physically (Cout, kh, kw, Cin), the
    K-major operand the GEMMs read through TMA (architectures/layers.py).  `load_state_dict(assign=True)`, `p.data = t` or a
    `memory_format` conversion can silently replace the storage by an NCHW-contiguous one; the kernels would then read the
    taps and channels in the wrong order.  Fail loudly instead."""
    if w.dim() == 4 and not w.permute(0, 2, 3, 1).is_contiguous():
        raise RuntimeError('convolution weight of shape {} has strides {}: the B200 kernels need channels-last storage '
                           '(p.data = p.data.contiguous(memory_format=torch.channels_last)); load_state_dict() without '
                           'assign=True keeps it'.format(tuple(w.shape), w.stride()))
    if w.dtype != torch.float32:
        raise RuntimeError('convolution weight has dtype {}, the B200 kernels compute on fp32 storage'.format(w.dtype))


# ====================================================================================== nodes
class ConvNode(object):
    """y = [relu]( conv(x, W) * scale + shift [+ residual] ); scale/shift = folded eval-mode BN or bias."""

    def __init__(self, x, y, conv, bn, residual, relu, scale, geom, col_src=None):
        self.x, self.y, self.conv, self.bn, self.residual, self.relu, self.scale = x, y, conv, bn, residual, relu, scale
        self.geom = geom          # (cout, kh, kw, cin, stride, pad, dil)
        self.col_src = col_src    # stem: (im2col matrix Act, padded K) kept for the weight gradient
        self.stem_input = None    # stem, input gradient requested: the image Act and the zero-padded (cout, 1, kpad) weights
        self.stem_wpad = None

    def backward(self, tape):
        K = tape.K
        cout, kh, kw, cin, stride, pad, dil = self.geom
        g = out_grad(self.y)
        if g is None:
            if self.residual is not None:
                tape.skip(self.residual)
            if self.col_src is None:
                tape.skip(self.x)
            elif self.stem_input is not None and self.stem_input.needs_grad:
                tape.skip(self.stem_input)
            return
        if self.y.parent is not None and self.relu:
            K.relu_gate(g, self.y)                 # slice outputs are gated here (parents carry plain sums)
        if self.residual is not None:
            tape.contribute_tensor(self.residual, g)
        bn = self.bn
        w = self.conv.weight
        bn_grads = bn is not None and tape.wants(bn.weight)
        st = None
        if bn_grads:
            st = getattr(self.y, 'fused_stats', None) if self.y.parent is None else None
            self.y.fused_stats = None
        if bn_grads and not tape.wants(w):
            # frozen convolution weights: recover xhat from the stored BN output
            dgam, acc = param_grad(bn.weight)
            dbet, acc2 = param_grad(bn.bias)
            assert acc == acc2
            if st is not None:
                K.bn_eval_param_grad_from_stats(st, bn.weight, bn.bias, dgam, dbet, acc)    # sums came with g
            else:
                K.bn_eval_param_grad(g, self.y, bn.weight, bn.bias, self.residual, dgam, dbet, acc)
            bn_grads = False
        if self.conv.bias is not None and tape.wants(self.conv.bias):
            db, acc = param_grad(self.conv.bias)
            K.colsum(g, db, acc)
        if self.col_src is not None:
            # stem: GEMM over the (recomputed) im2col matrix; weight gradient only
            if tape.wants(w):
                col, kpad = self.col_src        # the forward's column matrix is kept (0.67 GB at N=16, 512^2)
                dw_pad = K.empty((cout, 1, kpad), g.device)
                gflat = Act(g.base, 1, 1, g.rows, g.c, g.ld, g.off)
                K.conv_wgrad(gflat, col, dw_pad, cout, 1, 1, kpad, 1, 0, 1, row_scale=self.scale, accumulate=False)
                dw, acc = param_grad(w)
                K.copy_rows(dw, kh * kw * cin, dw_pad, kpad, cout, kh * kw * cin, acc)
                if bn_grads:
                    self._bn_grads_from_wgrad(K, st, g, w, dw, bn)
            x_img = self.stem_input
            if x_img is not None and x_img.needs_grad:
                # d/d(image) (VAT only): dcol = g . (scale * Wpad) as a 1x1 data-gradient GEMM, then the adjoint of im2col
                col, kpad = self.col_src
                wt, ldb = K.transpose_w(self.stem_wpad, cout, 1, kpad, scale=self.scale)
                dcol = col.like()
                gflat = Act(g.base, 1, 1, g.rows, g.c, g.ld, g.off)
                K.conv_dgrad(gflat, wt, kpad, 1, 1, cout, ldb, 1, 0, 1, dcol)
                oh, ow = self.y.h, self.y.w
                tape.contribute_kernel(x_img, lambda dst, accumulate, addend, gate, stats: K.col2im(
                    dcol, dst, kh, kw, stride, pad, dil, oh, ow, kpad, accumulate=accumulate), False)
            return
        if tape.wants(w):
            dw, acc = param_grad(w)
            K.conv_wgrad(g, self.x, dw, cout, kh, kw, cin, stride, pad, dil, row_scale=self.scale, accumulate=acc)
            if bn_grads:
                self._bn_grads_from_wgrad(K, st, g, w, dw, bn)
        if not self.x.needs_grad:
            return                                   # network input: no gradient needed
        wt, ldb = transposed_weights(K, self.conv, cout, kh * kw, cin, self.scale)

        def launch(dst, accumulate, addend, gate, stats_sub):
            return K.conv_dgrad(g, wt, cin, kh, kw, cout, ldb, stride, pad, dil, dst, addend=addend, gate=gate,
                                accumulate=accumulate, want_stats=stats_sub is not None,
                                stats_sub=stats_sub if stats_sub else None)
        tape.contribute_kernel(self.x, launch, fusable=(stride == 1))


    @staticmethod
    def _bn_grads_from_wgrad(K, st, g, w, dw, bn):
        """Frozen BN with trainable affine: dbeta += sum_pix g, dgamma = <W, dW>/gamma - invstd*mean*dbeta, where dW
        (already scaled by gamma*invstd and accumulated over passes by the wgrad epilogue) replaces every pass over
        the activations (see b2_bn_eval_param_grad_wdot in include/b200seg.h)."""
        dgam, acc = param_grad(bn.weight)
        dbet, acc2 = param_grad(bn.bias)
        assert acc == acc2
        K.bn_eval_param_grad_wdot(st, g, w, dw, bn, dgam, dbet, acc)


# train-mode BatchNorm backward without a residual recomputes its ReLU gate from the raw input instead of reading the stored
# output (b2_bn_bwd gate_beta); B200SEG_BN_REGATE=0 reads y as before (A/B timing)
REGATE = os.environ.get('B200SEG_BN_REGATE', '1') != '0'


class BNTrainNode(object):
    """y = dropout( relu( bn_train(raw) [+ residual] ) ); `parts`: [(first image, images, mean, rstd, dropmask)] -- one entry, or one
    per mini-batch of a multi-batch pass (each normalised with its own statistics)."""

    def __init__(self, raw, y, bn, parts, residual, relu, drop_scale):
        self.raw, self.y, self.bn, self.parts = raw, y, bn, parts
        self.residual, self.relu, self.drop_scale = residual, relu, drop_scale

    def backward(self, tape):
        K = tape.K
        dy = out_grad(self.y)
        if dy is None:
            tape.skip(self.raw)
            if self.residual is not None:
                tape.skip(self.residual)
            return
        bn = self.bn
        dx = self.raw.like()
        g_out = self.y.like() if self.residual is not None else None
        if tape.wants(bn.weight):
            dgam, acc = param_grad(bn.weight)
            dbet, _ = param_grad(bn.bias)
        else:
            dgam = dbet = None
            acc = False
        dyd = dy
        if any(p[4] is not None for p in self.parts) and dy.ld != dy.c:
            dyd = dy.like()
            K.copy_act(dyd, dy)
        whole = len(self.parts) == 1
        for i, (n0, n, mean, rstd, dropmask) in enumerate(self.parts):
            def sl(a):
                return a if (whole or a is None) else a.batch_slice(n0, n)
            # no residual: the gate y > 0 is recomputed from the raw input (same roundings as bn_apply), y is not read
            K.bn_bwd(sl(dyd), sl(self.raw), sl(self.y), mean, rstd, bn.weight, self.relu, dropmask, self.drop_scale, sl(dx),
                     dgam, dbet, acc or i > 0, g_out=sl(g_out),
                     gate_beta=bn.bias if (self.relu and self.residual is None and REGATE) else None)
        if self.raw.parent is not None:      # input = a channel prefix of a concatenation buffer (DenseNet norm1, train mode)
            tape.contribute_slice(self.raw, lambda dst, accumulate: K.copy_act(dst, dx, accumulate=accumulate))
        else:
            tape.contribute_tensor(self.raw, dx)
        if self.residual is not None:
            tape.contribute_tensor(self.residual, g_out)


class UpAddNode(object):
    """y = nearest x2 up-sampling of x [+ skip]  (DecoderBlock.forward, resunet.py:31-32; the final `up`, :87)."""

    def __init__(self, x, skip, y):
        self.x, self.skip, self.y = x, skip, y

    def backward(self, tape):
        K = tape.K
        dy = out_grad(self.y)
        if dy is None:
            tape.skip(self.x)
            if self.skip is not None:
                tape.skip(self.skip)
            return
        tape.contribute_kernel(self.x, lambda dst, accumulate, addend, gate, stats: K.upsample2x_bwd(
            dy, dst, accumulate=accumulate), False)
        if self.skip is not None:
            tape.contribute_tensor(self.skip, dy)           # d(skip) = dy: aliased, not copied


class MulMaskNode(object):
    """y = x * mask * scale: nn.Dropout applied to a raw convolution output (resunet.py:88)."""

    def __init__(self, x, y, mask, scale):
        self.x, self.y, self.mask, self.scale = x, y, mask, scale

    def backward(self, tape):
        K = tape.K
        dy = out_grad(self.y)
        if dy is None:
            tape.skip(self.x)
            return
        mask, scale = self.mask, self.scale

        def launch(dst, accumulate, addend, gate, stats):
            assert not accumulate, 'the dropped tensor has a single consumer'
            K.mul_mask(dy, mask, scale, dst)
        tape.contribute_kernel(self.x, launch, False)


class BNEvalNode(object):
    """y = [relu]( x * scale + shift ): a stand-alone eval-mode BatchNorm (DenseNet's pre-activation norm1 / transition norm /
    norm5 act on a concatenation, so they cannot be folded into the convolution that produced their input)."""

    def __init__(self, x, y, bn, scale):
        self.x, self.y, self.bn, self.scale = x, y, bn, scale

    def backward(self, tape):
        K = tape.K
        g = out_grad(self.y)                      # already gated by y > 0 (gate_on_grad)
        if g is None:
            tape.skip(self.x)
            return
        bn = self.bn
        if tape.wants(bn.weight):
            dgam, acc = param_grad(bn.weight)
            dbet, acc2 = param_grad(bn.bias)
            assert acc == acc2
            K.bn_eval_param_grad(g, self.y, bn.weight, bn.bias, None, dgam, dbet, acc)
        scale = self.scale
        tape.contribute_slice(self.x, lambda dst, accumulate: K.scale_channels(g, scale, dst, accumulate=accumulate))


class AvgPoolNode(object):
    def __init__(self, x, y):
        self.x, self.y = x, y

    def backward(self, tape):
        K = tape.K
        dy = out_grad(self.y)
        if dy is None:
            tape.skip(self.x)
            return
        tape.contribute_kernel(self.x, lambda dst, accumulate, addend, gate, stats: K.avgpool2x2_bwd(
            dy, dst, accumulate=accumulate), False)


class CopyNode(object):
    """dst (a slice of a concatenation buffer) = src."""

    def __init__(self, x, y):
        self.x, self.y = x, y

    def backward(self, tape):
        dy = out_grad(self.y)
        if dy is None:
            tape.skip(self.x)
            return
        tmp = self.x.like()
        tape.K.copy_act(tmp, dy)
        tape.contribute_tensor(self.x, tmp)


class MaxPoolNode(object):
    def __init__(self, x, y, idx):
        self.x, self.y, self.idx = x, y, idx

    def backward(self, tape):
        dy = out_grad(self.y)
        if dy is None:
            tape.skip(self.x)
            return
        tmp = self.x.like()
        tape.K.maxpool_bwd(dy, self.idx, tmp)
        tape.contribute_tensor(self.x, tmp)


class BilinearNode(object):
    def __init__(self, x, y, align_corners):
        self.x, self.y, self.align = x, y, align_corners

    def backward(self, tape):
        dy = out_grad(self.y)
        if dy is None:
            tape.skip(self.x)
            return
        K, align = tape.K, self.align
        tape.contribute_kernel(self.x, lambda dst, acc, addend, gate, stats: K.bilinear_bwd(dy, dst, align, accumulate=acc), False)


class GapNode(object):
    def __init__(self, x, y):
        self.x, self.y = x, y

    def backward(self, tape):
        dy = out_grad(self.y)
        if dy is None:
            tape.skip(self.x)
            return
        K = tape.K
        tape.contribute_kernel(self.x, lambda dst, acc, addend, gate, stats: K.gap_bwd(dy, dst, accumulate=acc), False)


class BcastNode(object):
    def __init__(self, v, y):
        self.v, self.y = v, y

    def backward(self, tape):
        dy = out_grad(self.y)
        if dy is None:
            tape.skip(self.v)
            return
        dv = self.v.like()
        tape.K.bcast_bwd(dy, dv)
        tape.contribute_tensor(self.v, dv)


class BatchSplitNode(object):
    """parts[i] = x[n_i : n_i + sizes[i]] (views).  Backward: the parts' gradients are batch slices of ONE buffer
    (their `grad_home`), which is handed to x as a single contribution -- no concatenation copy."""

    def __init__(self, x, parts, gbuf):
        self.x, self.parts, self.gbuf = x, parts, gbuf

    def backward(self, tape):
        K = tape.K
        any_grad = False
        for p in self.parts:
            if p.grad is None:
                K.fill_act(p.grad_home, 0.0)             # nobody consumed this part
            else:
                any_grad = True
                if p.grad is not p.grad_home:
                    K.copy_act(p.grad_home, p.grad)
            p.grad = None
            p.grad_home = None
            p.node = None
        if not any_grad:
            tape.skip(self.x)
            return
        tape.contribute_tensor(self.x, self.gbuf)


# ====================================================================================== forward ops
def batch_split(tape, x, sizes, delegate_gate=False):
    """Split a root activation along the batch dimension into independent root activations (views).  Used to run the
    batch-invariant (frozen-BN) part of a network ONCE over several mini-batches and the batch-dependent head
    (train-mode BatchNorm) per mini-batch.  `delegate_gate`: x has no other consumer, so the ReLU gate of its gradient
    is applied by the epilogues that finish the parts' gradients instead of a separate pass over x.grad."""
    assert x.parent is None and sum(sizes) == x.n
    parts, n0 = [], 0
    for n in sizes:
        parts.append(x.batch_slice(n0, n, name=x.name + '[{}:{}]'.format(n0, n0 + n)))
        n0 += n
    if not tape.enabled:
        return parts
    gbuf = x.like()
    n0 = 0
    for p, n in zip(parts, sizes):
        p.grad_home = gbuf.batch_slice(n0, n)
        n0 += n
    if delegate_gate and x.gate_on_grad:
        assert x.pending == 0, 'gate delegation needs the split to be the only consumer'
        x.gate_on_grad = False
        for p in parts:
            p.gate_on_grad = True
    node = BatchSplitNode(x, parts, gbuf)
    for p in parts:
        p.node = node
    tape.record(node, [x])
    return parts


def _conv_out_hw(h, w, k, stride, pad, dil):
    return ((h + 2 * pad - dil * (k - 1) - 1) // stride + 1, (w + 2 * pad - dil * (k - 1) - 1) // stride + 1)


def _geom(conv):
    cout, cin, kh, kw = conv.weight.shape
    return cout, kh, kw, cin, conv.stride, conv.padding, conv.dilation


def fold_bn(tape, bn):
    """scale/shift of an eval-mode BatchNorm (running statistics)."""
    key = (_GENERATION[0], tape.K.name, _versions(bn.weight, bn.bias, bn.running_mean, bn.running_var))
    hit = _FOLD_CACHE.get(bn)
    if hit is not None and hit[0] == key:
        return hit[1], hit[2]
    scale = torch.empty_like(bn.running_mean)
    shift = torch.empty_like(bn.running_mean)
    tape.K.bn_fold(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, scale, shift)
    _FOLD_CACHE[bn] = (key, scale, shift)
    return scale, shift


def transposed_weights(K, conv, cout, taps, cin, scale):
    """(cin, taps, pad4(cout)) dgrad operand with the folded BN scale, shared by the backward passes of one iteration."""
    w = conv.weight
    key = (_GENERATION[0], K.name, getattr(K, 'n_split', 1), _versions(w), None if scale is None else _versions(scale))
    hit = _WT_CACHE.get(conv)
    if hit is not None and hit[0] == key:
        return hit[1], hit[2]
    wt, ldb = K.transpose_w(w, cout, taps, cin, scale=scale)
    _WT_CACHE[conv] = (key, wt, ldb, scale)        # holds `scale` so its data_ptr cannot be recycled while cached
    return wt, ldb


def conv_bn_act(tape, x, conv, bn=None, residual=None, relu=False, out=None, dropout=None, ld_out=None):
    """conv -> (BatchNorm) -> (+residual) -> (ReLU) -> (Dropout), dispatching on the BN module's mode:
    eval-mode BN (frozen) is folded into the GEMM epilogue, train-mode BN runs statistics + apply."""
    K = tape.K
    cout, kh, kw, cin, stride, pad, dil = _geom(conv)
    assert x.c == cin, 'channel mismatch: {} vs {}'.format(x.c, cin)
    oh, ow = _conv_out_hw(x.h, x.w, kh, stride, pad, dil)
    w = conv.weight
    check_weight_layout(w)
    bias = conv.bias
    train_bn = bn is not None and bn.training
    if not train_bn:
        y = out if out is not None else Act.alloc(x.n, oh, ow, cout, x.device, ld=ld_out)
        scale = shift = None
        if bn is not None:
            scale, shift = fold_bn(tape, bn)
            assert bias is None
        elif bias is not None:
            shift = bias
        K.conv_fwd(x, w, cout, kh, kw, cin, cin, stride, pad, dil, y, scale=scale, shift=shift, addend=residual, relu=relu)
        y.gate_on_grad = bool(relu) and y.parent is None
        node = ConvNode(x, y, conv, bn, residual, relu, scale, (cout, kh, kw, cin, stride, pad, dil))
        y.node = node
        tape.record(node, [x] + ([residual] if residual is not None else []))
        assert dropout is None or not dropout.training, 'dropout after a frozen BN is not used by the reference nets'
        return y
    # train-mode BN: raw conv output -> batch statistics -> normalise (+residual, relu, dropout)
    raw = Act.alloc(x.n, oh, ow, cout, x.device)
    K.conv_fwd(x, w, cout, kh, kw, cin, cin, stride, pad, dil, raw, shift=bias)
    cnode = ConvNode(x, raw, conv, None, None, False, None, (cout, kh, kw, cin, stride, pad, dil))
    raw.node = cnode
    tape.record(cnode, [x])
    return bn_train(tape, raw, bn, residual=residual, relu=relu, out=out, dropout=dropout, ld_out=ld_out)


def bn_train(tape, raw, bn, residual=None, relu=False, out=None, dropout=None, ld_out=None):
    """Train-mode BatchNorm (+residual, ReLU, dropout).  In a multi-batch pass (tape.parts) every mini-batch is normalised with its
    own batch statistics, running statistics and dropout draws advance mini-batch by mini-batch -- exactly as in consecutive
    forward calls (reference: one `net(x)` call per mini-batch), while the convolutions around it run once over all of them."""
    K = tape.K
    momentum = bn.momentum if bn.momentum is not None else 0.1
    y = out if out is not None else Act.alloc(raw.n, raw.h, raw.w, raw.c, raw.device, ld=ld_out)
    sizes = tape.parts if (tape.parts is not None and len(tape.parts) > 1 and sum(tape.parts) == raw.n) else [raw.n]
    whole = len(sizes) == 1
    use_drop = dropout is not None and dropout.training and dropout.p > 0
    drop_scale = 1.0 / (1.0 - dropout.p) if use_drop else 1.0
    parts, n0 = [], 0
    for n in sizes:
        def sl(a):
            return a if (whole or a is None) else _batch_view(a, n0, n)
        mean = torch.empty_like(bn.running_mean)
        rstd = torch.empty_like(bn.running_mean)
        K.bn_stats(sl(raw), bn.eps, momentum, mean, rstd, bn.running_mean, bn.running_var)
        bn.num_batches_tracked += 1
        dropmask = dropout.next_mask(K, n, raw.h, raw.w, raw.c, raw.device) if use_drop else None
        K.bn_apply(sl(raw), mean, rstd, bn.weight, bn.bias, relu, dropmask, drop_scale, sl(y), residual=sl(residual))
        parts.append((n0, n, mean, rstd, dropmask))
        n0 += n
    node = BNTrainNode(raw, y, bn, parts, residual, relu, drop_scale)
    y.node = node
    tape.record(node, [raw] + ([residual] if residual is not None else []))
    return y


def _batch_view(a, n0, n):
    """Images [n0, n0 + n) of an activation (root or channel slice) as a plain view for a kernel call."""
    return Act(a.base[n0:n0 + n], n, a.h, a.w, a.c, a.ld, a.off, a.name)


def stem_conv(tape, x_nhwc, conv, bn):
    """Cin=3 7x7/s2 stem: explicit im2col (K = 147 padded to 160) + the same tensor-core GEMM."""
    K = tape.K
    cout, kh, kw, cin, stride, pad, dil = _geom(conv)
    oh, ow = _conv_out_hw(x_nhwc.h, x_nhwc.w, kh, stride, pad, dil)
    kreal = kh * kw * cin
    kpad = (kreal + 31) // 32 * 32
    check_weight_layout(conv.weight)
    col = K.im2col(x_nhwc, kh, kw, stride, pad, dil, oh, ow, kpad)
    wpad = torch.zeros((cout, 1, kpad), device=x_nhwc.device, dtype=torch.float32)
    K.copy_rows(wpad, kpad, conv.weight, kreal, cout, kreal, False)
    train_bn = bn.training
    tgt = Act.alloc(x_nhwc.n, oh, ow, cout, x_nhwc.device)
    flat = Act(tgt.base, 1, 1, tgt.rows, cout, cout, 0)
    if not train_bn:
        scale, shift = fold_bn(tape, bn)
        K.conv_fwd(col, wpad, cout, 1, 1, kpad, kpad, 1, 0, 1, flat, scale=scale, shift=shift, relu=True)
        tgt.gate_on_grad = True
        node = ConvNode(col, tgt, conv, bn, None, True, scale, (cout, kh, kw, cin, stride, pad, dil),
                        col_src=(col, kpad))
        tgt.node = node
        if x_nhwc.needs_grad:          # VAT: the gradient w.r.t. the image is wanted (netbase.b2_forward(input_grad=True))
            node.stem_input, node.stem_wpad = x_nhwc, wpad
            tape.record(node, [x_nhwc])
        else:
            tape.record(node, [])
        return tgt
    if x_nhwc.needs_grad:
        raise NotImplementedError('input gradient through a train-mode BatchNorm stem (the reference computes the VAT '
                                  'direction in eval mode, train_seg_semisup_vat_mt.py:237)')
    K.conv_fwd(col, wpad, cout, 1, 1, kpad, kpad, 1, 0, 1, flat)
    cnode = ConvNode(col, tgt, conv, None, None, False, None, (cout, kh, kw, cin, stride, pad, dil),
                     col_src=(col, kpad))
    tgt.node = cnode
    tape.record(cnode, [])
    return bn_train(tape, tgt, bn, relu=True)


def upsample2x_add(tape, x, skip=None):
    """nn.Upsample(scale_factor=2) (nearest) followed by the skip addition of the U-Net decoders."""
    if skip is not None:
        assert (skip.n, skip.h, skip.w, skip.c) == (x.n, 2 * x.h, 2 * x.w, x.c), 'skip connection shape mismatch'
    y = Act.alloc(x.n, 2 * x.h, 2 * x.w, x.c, x.device)
    tape.K.upsample2x_add(x, skip, y)
    node = UpAddNode(x, skip, y)
    y.node = node
    tape.record(node, [x] + ([skip] if skip is not None else []))
    return y


def dropout_raw(tape, x, dropout):
    """Dropout on a tensor that is not the output of a fused BN/ReLU (train mode only; identity otherwise)."""
    if dropout is None or not dropout.training or dropout.p <= 0:
        return x
    mask = dropout.next_mask(tape.K, x.n, x.h, x.w, x.c, x.device)
    scale = 1.0 / (1.0 - dropout.p)
    y = x.like()
    tape.K.mul_mask(x, mask, scale, y)
    node = MulMaskNode(x, y, mask, scale)
    y.node = node
    tape.record(node, [x])
    return y


_IDENTITY_STATS = {}


def bn_eval_act(tape, x, bn, relu=True):
    """Stand-alone BatchNorm (+ReLU) of an activation or of a channel slice of one (DenseNet's pre-activation norms); returns a
    dense activation.  Eval mode: folded scale / shift in one pass; train mode: the statistics + apply kernels."""
    if bn.training:          # batch statistics over the (slice of the) concatenation: the train-mode BN path, dense output
        return bn_train(tape, x, bn, relu=relu)
    K = tape.K
    scale, shift = fold_bn(tape, bn)
    key = (K.name, x.c, str(x.device))
    if key not in _IDENTITY_STATS:          # (x - 0) * 1 * scale + shift through the fused BN-apply kernel
        _IDENTITY_STATS[key] = (torch.zeros(x.c, device=x.device), torch.ones(x.c, device=x.device))
    zeros, ones = _IDENTITY_STATS[key]
    y = Act.alloc(x.n, x.h, x.w, x.c, x.device)
    K.bn_apply(x, zeros, ones, scale, shift, relu, None, 1.0, y)
    y.gate_on_grad = bool(relu)
    node = BNEvalNode(x, y, bn, scale)
    y.node = node
    tape.record(node, [x])
    return y


def avgpool2x2(tape, x, out=None):
    y = out if out is not None else Act.alloc(x.n, x.h // 2, x.w // 2, x.c, x.device)
    assert (y.n, y.h, y.w, y.c) == (x.n, x.h // 2, x.w // 2, x.c)
    tape.K.avgpool2x2(x, y)
    node = AvgPoolNode(x, y)
    y.node = node
    tape.record(node, [x])
    return y


def copy_into(tape, x, out):
    """out (a channel slice of a concatenation buffer) = x."""
    assert (out.n, out.h, out.w, out.c) == (x.n, x.h, x.w, x.c)
    tape.K.copy_act(out, x)
    node = CopyNode(x, out)
    out.node = node
    tape.record(node, [x])
    return out


def maxpool3x3s2(tape, x, ceil_mode):
    K = tape.K
    if ceil_mode:
        oh, ow = -(-(x.h + 2 - 3) // 2) + 1, -(-(x.w + 2 - 3) // 2) + 1
        if (oh - 1) * 2 >= x.h + 1:     # PyTorch: the last window must start inside the input or left padding
            oh -= 1
        if (ow - 1) * 2 >= x.w + 1:
            ow -= 1
    else:
        oh, ow = (x.h + 2 - 3) // 2 + 1, (x.w + 2 - 3) // 2 + 1
    y = Act.alloc(x.n, oh, ow, x.c, x.device)
    idx = torch.empty((x.n, oh, ow, x.c), device=x.device, dtype=torch.uint8)
    xin = x
    if x.ld != x.c:
        xin = x.like()
        K.copy_act(xin, x)
    K.maxpool_fwd(xin, y, idx)
    node = MaxPoolNode(x, y, idx)
    y.node = node
    tape.record(node, [x])
    return y


def bilinear(tape, x, oh, ow, align_corners, out=None):
    y = out if out is not None else Act.alloc(x.n, oh, ow, x.c, x.device)
    tape.K.bilinear_fwd(x, y, align_corners)
    node = BilinearNode(x, y, align_corners)
    y.node = node
    tape.record(node, [x])
    return y


def global_avg_pool(tape, x):
    y = Act.alloc(x.n, 1, 1, x.c, x.device, ld=x.c)     # dense (N, C) vector
    tape.K.gap_fwd(x, y)
    node = GapNode(x, y)
    y.node = node
    tape.record(node, [x])
    return y


def broadcast(tape, v, out):
    tape.K.bcast_fwd(v, out)
    node = BcastNode(v, out)
    out.node = node
    tape.record(node, [v])
    return out


def to_logits_nchw(tape, x, out_h, out_w, align_corners):
    """Final bilinear resize to the input resolution, written as an NCHW tensor (the reference's
    output layout).  Returns (logits, backward_entry)."""
    logits = torch.empty((x.n, x.c, out_h, out_w), device=x.device, dtype=torch.float32)
    tape.K.bilinear_fwd_nchw(x, logits, align_corners)
    if tape.enabled:
        root = x
        while root.parent is not None:
            root = root.parent
        root.pending += 1
    return logits


def seed_output_grad(tape, x, dlogits, align_corners, scale_dev=None, scale_host=1.0):
    """Start of the backward pass: d(loss)/d(low-res logits) from d(loss)/d(logits) (NCHW)."""
    K = tape.K
    tape.contribute_kernel(x, lambda dst, acc, addend, gate, stats: K.bilinear_bwd_nchw(
        dlogits, dst, align_corners, scale_dev=scale_dev, scale_host=scale_host, accumulate=acc), False)

"""One iteration of the CutMix / CutOut mean-teacher loop on the B200 kernels -- and of its sibling scripts, which share
everything but the unsupervised branch: ICT (`unsupervised_ict`), augmentation consistency (`unsupervised_aug`), VAT
(`unsupervised_vat`).

This is the body of the reference's training loop (train_seg_semisup_mask_mt.py:287-476) with the
same order of operations and the same arithmetic, restructured for the GPU:

  reference (per iteration)                          here
  ------------------------------------------------   ---------------------------------------------
  CE = CrossEntropyLoss(ignore 255); backward        one fused CE kernel (loss + dlogits), engine tape
  x0*(1-m)+x1*m for image and valid mask (6 ATen)    b2_mix (x2 launches)
  2 teacher forwards, 1 student forward              engine forwards (teacher without a tape)
  logit mix, 2 softmax, max, >=, mean, (p-q)^2 ...   ONE fused consistency kernel (loss, conf rate, dlogits)
  float(conf_mask.mean()), float(loss) host syncs    scalars stay on the device (read them when logging)
  loss.backward() through autograd                   engine tape; the 1/N, conf-rate, ramp and weight scalars
                                                     are applied by the first backward kernel from a device scalar
  student_optim.step(); teacher_optim.step()         ONE fused optimiser + EMA launch (optim.FusedOptimizer); torch.optim +
                                                     the multi-tensor EMA kernel when an ordinary optimiser is passed
  (single GPU)                                       optional data parallelism: ONE all-reduce(avg) of the
                                                     flat student-gradient buffer before the optimiser step
"""
import torch

from . import ops as O


class FlatGrads(object):
    """All student gradients as views into one flat fp32 buffer: a single fill kernel zeroes them and a
    single NCCL all-reduce averages them across ranks."""

    def __init__(self, params):
        seen, uniq = set(), []
        for p in params:
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                uniq.append(p)
        self.params = uniq
        total = sum(p.numel() for p in uniq)
        dev = uniq[0].device
        self.flat = torch.zeros((total,), device=dev, dtype=torch.float32)
        off = 0
        for p in uniq:
            n = p.numel()
            view = torch.as_strided(self.flat, p.shape, p.stride(), off)   # same (channels-last) layout as p
            p.grad = view
            off += n
        self.numel = total

    def attach(self):
        """Re-attach the views (after `optimizer.zero_grad(set_to_none=True)` removed them)."""
        off = 0
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + 4 * off:
                p.grad = torch.as_strided(self.flat, p.shape, p.stride(), off)
            off += p.numel()

    def zero(self, backend):
        self.attach()
        backend.fill(self.flat, 0.0)

    def buckets(self, n_buckets):
        """The flat buffer cut into `n_buckets` contiguous ranges at parameter boundaries, LAST range first (parameters are laid
        out in forward order, the backward pass finishes them back to front): [(first element, one-past-last element,
        [parameters])].  Sizes shrink towards the front (each bucket takes half of what is left, the final one the rest): the
        bucket that is ready last -- the first layers, whose all-reduce cannot hide under any remaining kernel -- is the
        smallest."""
        total = self.numel
        out, params, hi, off = [], [], total, total
        target = max(1, total // 2)
        for p in reversed(self.params):
            off -= p.numel()
            params.append(p)
            if hi - off >= target and len(out) < n_buckets - 1:
                out.append((off, hi, params))
                params, hi = [], off
                target = max(1, off // 2)
        if hi > 0 or params:
            out.append((0, hi, params))
        return out


def make_optimizer(student_net, opt_type, learning_rate, sgd_momentum=0.9, sgd_nesterov=False, sgd_weight_decay=5e-4,
                   capturable=False, fused_kernel=False):
    """torch.optim optimiser on the reference's parameter groups (train_seg_semisup_mask_mt.py:90-100): group 0 =
    `pretrained_parameters()` at 0.1 x lr, group 1 = `new_parameters()` at lr.  DeepLab v2's group 0 repeats tensors
    (reference quirk): a tensor listed k times must receive k sequential updates per step, which only the
    per-tensor implementation (`foreach=False`, what torch 1.4 did) reproduces.  Without duplicates the single
    fused multi-tensor kernel of torch is the same algorithm and is used instead."""
    import warnings
    g0 = list(student_net.pretrained_parameters())
    g1 = list(student_net.new_parameters())
    if fused_kernel:
        # ONE sm_100a launch for the optimiser step (+ the teacher's EMA step when MeanTeacherStep pairs them), with the
        # k-sequential-updates rule for duplicated entries implemented in the kernel (cutmix_semisup_seg_b200/optim.py)
        from .optim import FusedOptimizer
        groups = [dict(params=g0, lr=learning_rate * 0.1), dict(params=g1, lr=learning_rate)]
        if opt_type == 'adam':
            return FusedOptimizer(groups, 'adam', lr=learning_rate)
        if opt_type == 'sgd':
            return FusedOptimizer(groups, 'sgd', lr=learning_rate, momentum=sgd_momentum, nesterov=sgd_nesterov,
                                  weight_decay=sgd_weight_decay)
        raise ValueError('Unknown opt_type {}'.format(opt_type))
    dup = len(set(id(p) for p in g0 + g1)) != len(g0) + len(g1)
    groups = [dict(params=g0, lr=learning_rate * 0.1), dict(params=g1, lr=learning_rate)]
    on_cuda = all(p.is_cuda for p in g0 + g1)
    kw = dict(foreach=False) if (dup or not on_cuda) else dict(fused=True)
    if capturable and opt_type == 'adam' and on_cuda:
        kw['capturable'] = True          # step counters live on the device: the step can be replayed from a CUDA graph
        # ... and so must the learning rates: a python float would be baked into the captured kernels and every LR schedule
        # silently ignored on replay.  torch's schedulers update tensor learning rates in place (`fill_`).
        dev = (g0 + g1)[0].device
        for g in groups:
            g['lr'] = torch.tensor(float(g['lr']), device=dev, dtype=torch.float32)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        if opt_type == 'adam':
            opt = torch.optim.Adam(groups, **kw)
        elif opt_type == 'sgd':
            opt = torch.optim.SGD(groups, momentum=sgd_momentum, nesterov=sgd_nesterov, weight_decay=sgd_weight_decay, **kw)
        else:
            raise ValueError('Unknown opt_type {}'.format(opt_type))
    opt.b2_note = '{} ({})'.format(opt_type, 'per-tensor, duplicated reference group' if dup else 'torch fused multi-tensor')
    return opt


def average_gradients(flat, dist, group=None):
    """ONE collective per iteration: average the flat student-gradient buffer over the data-parallel ranks
    (natural shard: every rank holds the gradient of its own mini-batch).  NCCL reduces with ncclAvg over
    NVLink/NVSwitch; other backends (gloo, used by the CPU tests of this host logic) sum and divide."""
    world = dist.get_world_size(group)
    if world == 1:
        return
    if dist.get_backend(group) == 'nccl':
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)


class BucketedAllReduce(object):
    """The iteration's ONE exchange step (average of the flat student-gradient buffer over the data-parallel ranks, SURVEY.md
    8e), issued as a few bucket all-reduces WHILE the backward pass is still running: a bucket (a contiguous range at the end of
    the flat buffer = the layers whose gradients are final first) is handed to NCCL as soon as the backward pass has produced
    it; NCCL runs on its own stream, ordered after the producing kernels, under the remaining backward kernels.  Rank skew and
    the transfer itself are hidden up to the last bucket (237 MB over NVLink is ~0.6 ms of an ~147 ms iteration; the exposed
    single all-reduce cost ~3.5 ms at 8 GPUs including skew, SCALE_r01).  `finish()` makes the current stream wait for all of
    them (before the optimiser step).  Other backends (gloo, CPU tests of this host logic): the same calls, summed and divided."""

    def __init__(self, flat, dist, group, n_buckets=4):
        self.flat, self.dist, self.group = flat, dist, group
        self.world = dist.get_world_size(group)
        self.nccl = dist.get_backend(group) == 'nccl'
        self.spans = flat.buckets(n_buckets)
        self.pending = []
        self.launched = 0

    def ready_cuts(self, ready):
        """{tape node index: launch bucket b}: bucket b may go once the earliest node owning one of its parameters has run."""
        cuts = {}
        for b, (lo, hi, params) in enumerate(self.spans):
            idx = [ready[id(p)] for p in params if id(p) in ready]
            at = min(idx) if idx else -1            # no layer of the pass uses these parameters: zero gradients, ready at once
            prev = cuts.get(at)
            cuts[at] = (lambda b=b, prev=prev: ((prev() if prev else None), self.launch(b)))
        return cuts

    def launch(self, b):
        lo, hi, _ = self.spans[b]
        view = self.flat.flat[lo:hi]
        if self.nccl:
            work = self.dist.all_reduce(view, op=self.dist.ReduceOp.AVG, group=self.group, async_op=True)
        else:
            work = self.dist.all_reduce(view, op=self.dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.pending.append((work, view))
        self.launched += 1

    def finish(self):
        """All buckets launched so far are complete as far as the current stream is concerned."""
        for work, view in self.pending:
            work.wait()
            if not self.nccl:
                view.div_(self.world)
        done, self.pending = len(self.pending), []
        return done


class MeanTeacherStep(object):
    def __init__(self, student_net, teacher_net, student_optim, teacher_optim, mask_generator, cons_loss_fn='var',
                 cons_weight=1.0, conf_thresh=0.97, conf_per_pixel=False, rampup=-1, mask_mix=True,
                 unsup_batch_ratio=1, dist_group=None, use_flat_grads=True, use_cuda_graph=False, batch_trunk=True,
                 vat_radius=0.5, adaptive_vat_radius=False, vat_dir_from_student=False):
        self.student_net, self.teacher_net = student_net, teacher_net
        self.student_optim, self.teacher_optim = student_optim, teacher_optim
        self.mask_generator = mask_generator
        self.cons_loss_fn, self.cons_weight = cons_loss_fn, cons_weight
        self.conf_thresh, self.conf_per_pixel = conf_thresh, conf_per_pixel
        self.rampup, self.mask_mix, self.unsup_batch_ratio = rampup, mask_mix, unsup_batch_ratio
        # VAT (train_seg_semisup_vat_mt.py:624-626); used only by batches that carry the key 'vat'
        self.vat_radius, self.adaptive_vat_radius, self.vat_dir_from_student = vat_radius, adaptive_vat_radius, vat_dir_from_student
        self.be = O.default_backend()
        self.optim_note = getattr(student_optim, 'b2_note', type(student_optim).__name__)
        self.world = 1
        self.dist = None
        if dist_group is not None:
            import torch.distributed as dist
            self.dist = dist
            self.group = dist_group if dist_group is not True else None
            self.world = dist.get_world_size(self.group)
        self.flat = FlatGrads(list(student_net.parameters())) if (use_flat_grads or self.world > 1) else None
        # gradient exchange overlapped with the backward pass (batched-trunk iterations: ONE backward pass produces every
        # gradient); B200SEG_GRAD_BUCKETS=1 restores the single all-reduce after the pass
        import os as _os
        self.grad_buckets = int(_os.environ.get('B200SEG_GRAD_BUCKETS', '4'))
        self._bucketed = BucketedAllReduce(self.flat, self.dist, self.group, self.grad_buckets) \
            if (self.world > 1 and self.grad_buckets > 1) else None
        self._overlapped = False         # this iteration's buckets were launched during the backward pass
        # CUDA-graph replay of the iteration (the eager path issues ~6000 launches per iteration from Python and is
        # host-bound: profiles/r01_v2_*).  Two graphs: forward/backward/losses, and optimiser + EMA, with the
        # gradient all-reduce between them.
        self.use_cuda_graph = use_cuda_graph
        # run the frozen-BN trunk once per network and iteration over the concatenated mini-batches (see
        # netbase.B2SegNet.b2_forward_multi); falls back to the reference's pass-by-pass order when BatchNorm is not frozen
        self.batch_trunk = batch_trunk
        self._graph = None
        self._segmenter = None           # set while capturing: turns bucket-ready points into CUDA-graph boundaries
        self.launches_per_replay = 0

    # ------------------------------------------------------------------------------------------
    def _zero_grad(self):
        if self.flat is not None:
            self.flat.zero(self.be)
        else:
            self.student_optim.zero_grad()

    def _allreduce(self):
        if self.world > 1:
            if self._overlapped:                     # launched bucket by bucket under the backward pass: wait for them
                self._bucketed.finish()
                self._overlapped = False
            else:
                average_gradients(self.flat.flat, self.dist, self.group)

    def supervised(self, batch_x, batch_y):
        """Lines 296-301: student forward, CE(ignore 255), backward.  Returns the loss as a device scalar."""
        logits, state = self.student_net.b2_forward(batch_x, record=True)
        labels = batch_y[:, 0] if batch_y.dim() == 4 else batch_y
        out3, dlogits = self.be.cross_entropy(logits, labels.contiguous(), ignore_index=255)
        self.student_net.b2_backward(state, dlogits, scale_dev=out3[2:3])
        return out3[0]

    def unsupervised_mix(self, ux0_tea, ux0_stu, um0, ux1_tea, ux1_stu, um1, mask_params, ramp_val=1.0):
        """Lines 309-369 + 406-459 (mix mode)."""
        be = self.be
        masks = self.mask_generator.torch_masks_from_params(mask_params, ux0_stu.shape[2:4], ux0_stu.device)
        masks = masks.contiguous()
        ux_mixed = be.mix(ux0_stu, ux1_stu, masks)                    # :350
        um_mixed = be.mix(um0, um1, masks)                            # :351
        with torch.no_grad():                                          # :354-356
            l0 = self.teacher_net.b2_forward(ux0_tea, record=False)[0]
            l1 = self.teacher_net.b2_forward(ux1_tea, record=False)[0]
        ls, state = self.student_net.b2_forward(ux_mixed, record=True)            # :358
        ramp = ramp_val if self.rampup > 0 else 1.0
        out4, dls = be.consistency(l0, l1, ls, masks, um_mixed, self.cons_loss_fn, self.conf_thresh, self.conf_per_pixel,
                                   ramp, self.cons_weight)
        self.student_net.b2_backward(state, dls, scale_dev=out4[2:3])             # :458-459
        return out4

    def unsupervised_ict(self, ux0_tea, ux0_stu, um0, ux1_tea, ux1_stu, um1, ict_mix_factors, ramp_val=1.0):
        """ICT (train_seg_semisup_ict.py:306-392): images and valid masks mixed with one Beta factor per sample, teacher
        probabilities / confidences mixed with the same factor inside the fused loss kernel."""
        be = self.be
        f = ict_mix_factors.reshape(-1).to(torch.float32)
        ux_mixed = be.mix_per_sample(ux0_stu, ux1_stu, f)             # ict :310
        um_mixed = be.mix_per_sample(um0, um1, f)                     # ict :311
        with torch.no_grad():                                          # ict :314-316
            l0 = self.teacher_net.b2_forward(ux0_tea, record=False)[0]
            l1 = self.teacher_net.b2_forward(ux1_tea, record=False)[0]
        ls, state = self.student_net.b2_forward(ux_mixed, record=True)            # ict :318
        ramp = ramp_val if self.rampup > 0 else 1.0
        out4, dls = be.ict_consistency(l0, l1, ls, f, um_mixed, self.cons_loss_fn, self.conf_thresh, self.conf_per_pixel,
                                       ramp, self.cons_weight)
        self.student_net.b2_backward(state, dls, scale_dev=out4[2:3])             # ict :389-390
        return out4

    def _check_aug_loss_fn(self):
        if self.cons_loss_fn == 'logits_var':
            # same failure as the reference: train_seg_semisup_aug_mt.py:373 reads `delta_prob`, which only the `var` branch
            # assigns, so the first unsupervised batch raises UnboundLocalError (a NameError)
            raise UnboundLocalError("local variable 'delta_prob' referenced before assignment "
                                    "(train_seg_semisup_aug_mt.py:373: the reference's logits_var branch cannot run)")

    def unsupervised_aug(self, ux0, um0, ux1, um1, xf0_to_1, ramp_val=1.0):
        """Augmentation-driven consistency (train_seg_semisup_aug_mt.py:275-402): teacher on view 0, student on view 1, the
        teacher's logits / probabilities / valid mask resampled into the student's frame under the affine map `xf0_to_1`
        (N,2,3) inside the fused loss kernel."""
        be = self.be
        self._check_aug_loss_fn()
        with torch.no_grad():                                          # aug :291-293
            lt = self.teacher_net.b2_forward(ux0, record=False)[0]
        ls, state = self.student_net.b2_forward(ux1, record=True)      # aug :295
        ramp = ramp_val if self.rampup > 0 else 1.0
        out4, dls = be.aug_consistency(lt, ls, xf0_to_1.to(torch.float32), um0, um1, self.cons_loss_fn, self.conf_thresh,
                                       self.conf_per_pixel, ramp, self.cons_weight)
        self.student_net.b2_backward(state, dls, scale_dev=out4[2:3])  # aug :397-398
        return out4

    VAT_LOSS_FNS = ('var', 'bce', 'kld', 'logits_var')        # train_seg_semisup_vat_mt.py:251-262

    def vat_perturbation(self, x, x_hat, noise=None):
        """`vat_perburbation` + `vat_direction` of the reference (train_seg_semisup_vat_mt.py:228-301): the adversarial
        perturbation (N,3,H,W) for the student's view `x_hat`, from the input gradient of the consistency loss between the
        direction network's predictions on `x` and on `x_hat + eps`.  `noise`: the N(0,1) draw behind eps (reference:
        torch.randn, :222); given explicitly by the parity tests.  Returns the perturbed images x_hat + r_adv (:392)."""
        be = self.be
        if self.cons_loss_fn not in self.VAT_LOSS_FNS:
            raise ValueError('Unknown consistency loss function {}'.format(self.cons_loss_fn))      # :261-262
        dir_net = self.student_net if self.vat_dir_from_student else self.teacher_net               # :102-105
        dir_net.eval()                 # :237 -- and, like the reference, it STAYS in eval mode until the next epoch's .train()
        with torch.no_grad():
            y_logits = dir_net.b2_forward(x, record=False)[0]                                       # :238-239
            n, c, h, w = x.shape
            noise_scale = 1.0e-6 * h * w / 1000                                                     # :243
            if noise is None:
                noise = torch.randn(x.shape, dtype=torch.float32, device=x.device)                  # :222
            x_eps = be.add_scaled_per_sample(x_hat, noise, be.sample_l2norm(noise), noise_scale)    # :223, :247
            e_logits, state = dir_net.b2_forward(x_eps, record=True, input_grad=True)               # :247
            # d(loss)/d(eps): the loss kernel's un-scaled gradient is that of the reference's SUMMED loss (:251-260; for
            # logits_var up to the constant 1/sqrt(C)); positive constants cancel in the normalisation below
            _, dle = be.consistency(y_logits, None, e_logits, None, None, self.cons_loss_fn, 0.0, False, 1.0, 1.0)
            eps_adv = dir_net.b2_backward(state, dle, param_grads=False)                            # :265-268
            if self.adaptive_vat_radius:
                radius = be.vat_adaptive_radius(x_hat, self.vat_radius)                             # :277-296
            else:
                radius = self.vat_radius * float(c * h * w) ** 0.5                                  # :298-299
            return be.add_scaled_per_sample(x_hat, eps_adv, be.sample_l2norm(eps_adv), radius)      # :271, :301, :392

    def unsupervised_vat(self, ux_tea, ux_stu, um, ramp_val=1.0, noise=None):
        """VAT consistency (train_seg_semisup_vat_mt.py:364-455): student on the adversarially perturbed image, teacher on
        the clean one, the confidence-thresholded consistency loss of the CutOut branch with the valid mask as loss mask."""
        be = self.be
        ux_adv = self.vat_perturbation(ux_tea, ux_stu, noise)                       # vat :389-392
        with torch.no_grad():
            lt = self.teacher_net.b2_forward(ux_tea, record=False)[0]               # vat :395-396
        ls, state = self.student_net.b2_forward(ux_adv, record=True)                # vat :398
        ramp = ramp_val if self.rampup > 0 else 1.0
        out4, dls = be.consistency(lt, None, ls, None, um, self.cons_loss_fn, self.conf_thresh, self.conf_per_pixel, ramp,
                                   self.cons_weight)                                # vat :400-448
        self.student_net.b2_backward(state, dls, scale_dev=out4[2:3])               # vat :451-452
        return out4

    def unsupervised_cut(self, ux_tea, ux_stu, um, mask_params, ramp_val=1.0):
        """Lines 371-401 + 406-459 (cut / CutOut mode)."""
        be = self.be
        masks = self.mask_generator.torch_masks_from_params(mask_params, ux_stu.shape[2:4], ux_stu.device).contiguous()
        ux_cut = be.mix(ux_stu, None, masks)                          # :389
        with torch.no_grad():
            lt = self.teacher_net.b2_forward(ux_tea, record=False)[0]             # :393
        ls, state = self.student_net.b2_forward(ux_cut, record=True)              # :395
        loss_mask = be.mix(um, None, masks)                           # :401  (um * mask)
        ramp = ramp_val if self.rampup > 0 else 1.0
        out4, dls = be.consistency(lt, None, ls, None, loss_mask, self.cons_loss_fn, self.conf_thresh,
                                   self.conf_per_pixel, ramp, self.cons_weight)
        self.student_net.b2_backward(state, dls, scale_dev=out4[2:3])
        return out4

    def _can_batch_trunk(self, unsup_batches):
        if any(ub is not None and 'vat' in ub for ub in unsup_batches):
            return False           # VAT changes the networks' train / eval mode in the middle of the iteration: pass by pass
        return (self.batch_trunk and self.cons_weight > 0.0 and len(unsup_batches) == 1 and
                self.teacher_net is not self.student_net and
                hasattr(self.student_net, 'b2_forward_multi') and hasattr(self.teacher_net, 'b2_forward_multi') and
                self.student_net.b2_trunk_is_batch_invariant() and self.teacher_net.b2_trunk_is_batch_invariant())

    def _fwd_bwd_batched(self, sup_batch, ub, ramp_val):
        """The same iteration with the batch-invariant trunks run once per network: student on [labelled ; mixed]
        images, teacher on [view 0 ; view 1].  Per-sample arithmetic, head BatchNorm statistics / dropout order and the
        losses are those of `supervised` + `unsupervised_mix` / `unsupervised_cut`; parameter gradients are summed by one
        backward pass instead of two accumulating ones (same sums, different fp32 association)."""
        be = self.be
        batch_x, batch_y = sup_batch
        ict = 'ict_mix_factors' in ub
        aug = 'xf0_to_1' in ub
        if aug:
            self._check_aug_loss_fn()
            ux_in = ub['ux1']                                              # aug :295 (no mixing: the views differ geometrically)
        elif ict:
            f = ub['ict_mix_factors'].reshape(-1).to(torch.float32)
            ux_in = be.mix_per_sample(ub['ux0_stu'], ub['ux1_stu'], f)    # ict :310
            loss_mask = be.mix_per_sample(ub['um0'], ub['um1'], f)        # ict :311
        elif self.mask_mix:
            ux_stu = ub['ux0_stu']
            masks = self.mask_generator.torch_masks_from_params(ub['mask_params'], ux_stu.shape[2:4], ux_stu.device).contiguous()
            ux_in = be.mix(ux_stu, ub['ux1_stu'], masks)                   # :350
            loss_mask = be.mix(ub['um0'], ub['um1'], masks)                # :351
        else:
            ux_stu = ub['ux_stu']
            masks = self.mask_generator.torch_masks_from_params(ub['mask_params'], ux_stu.shape[2:4], ux_stu.device).contiguous()
            ux_in = be.mix(ux_stu, None, masks)                            # :389
            loss_mask = be.mix(ub['um'], None, masks)                      # :401
        (sup_logits, ls), state = self.student_net.b2_forward_multi([batch_x, ux_in], record=True)   # :299, :358 / :395
        with torch.no_grad():                                              # :354-356 / :393
            if aug:
                l0, l1 = self.teacher_net.b2_forward(ub['ux0'], record=False)[0], None     # aug :291-293
            elif self.mask_mix or ict:
                (l0, l1), _ = self.teacher_net.b2_forward_multi([ub['ux0_tea'], ub['ux1_tea']], record=False)
            else:
                l0, l1 = self.teacher_net.b2_forward(ub['ux_tea'], record=False)[0], None
        labels = batch_y[:, 0] if batch_y.dim() == 4 else batch_y
        out3, dsup = be.cross_entropy(sup_logits, labels.contiguous(), ignore_index=255)        # :300
        ramp = ramp_val if self.rampup > 0 else 1.0
        if aug:
            out4, dls = be.aug_consistency(l0, ls, ub['xf0_to_1'].to(torch.float32), ub['um0'], ub['um1'], self.cons_loss_fn,
                                           self.conf_thresh, self.conf_per_pixel, ramp, self.cons_weight)
        elif ict:
            out4, dls = be.ict_consistency(l0, l1, ls, f, loss_mask, self.cons_loss_fn, self.conf_thresh, self.conf_per_pixel,
                                           ramp, self.cons_weight)
        else:
            out4, dls = be.consistency(l0, l1, ls, masks if self.mask_mix else None, loss_mask, self.cons_loss_fn,
                                       self.conf_thresh, self.conf_per_pixel, ramp, self.cons_weight)
        cuts_for = None
        if self._bucketed is not None and self._segmenter is None:
            cuts_for = self._bucketed.ready_cuts              # eager: bucket all-reduces launched from inside the pass
            self._overlapped = True
        elif self._bucketed is not None:
            cuts_for = self._segmenter                        # graph capture: the pass is cut into one graph per bucket
        self.student_net.b2_backward_multi(state, [dsup, dls], [out3[2:3], out4[2:3]], cuts_for=cuts_for)   # :301, :459
        return {'sup_loss': out3[0], 'cons_loss': out4[0], 'conf_rate': out4[1]}

    def _capture(self, sup_batch, unsup_batches, ramp_val):
        dev = sup_batch[0].device if sup_batch[0].is_cuda else next(self.student_net.parameters()).device
        st_sup = tuple(torch.empty(t.shape, dtype=t.dtype, device=dev) for t in sup_batch)
        st_uns = []
        for ub in unsup_batches:
            cache, d = {}, {}
            for k, v in ub.items():
                if id(v) not in cache:
                    cache[id(v)] = torch.empty(v.shape, dtype=v.dtype, device=dev)
                d[k] = cache[id(v)]
            st_uns.append(d)
        self._static = (st_sup, st_uns)
        self._staging, self._staged = None, None
        self._load_static(sup_batch, unsup_batches)
        snap = self._snapshot()            # the warm-up iteration below must not count as a training step
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):                      # warm-up on a side stream (allocator, lazy inits)
            self._fwd_bwd(st_sup, st_uns, ramp_val)
            self._allreduce()
            self._opt_ema()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        l0 = self.be.launches
        segs, plan = [], []
        if self._bucketed is not None and self._can_batch_trunk(st_uns):
            # Data parallel: the forward / backward capture is cut into one CUDA graph per gradient bucket, at the points of
            # the backward pass where a bucket's gradients are final; between two replays the bucket's all-reduce is handed to
            # NCCL (eagerly, on NCCL's stream), so it runs under the next segment's kernels.  Manual capture_begin / capture_end
            # (torch.cuda.graph cannot be left from inside the pass); all segments share one memory pool and replay in
            # capture order.
            import gc
            torch.cuda.synchronize(); gc.collect(); torch.cuda.empty_cache()
            cap = torch.cuda.Stream()
            cap.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(cap):
                pool = torch.cuda.graph_pool_handle()
                g = torch.cuda.CUDAGraph()
                g.capture_begin(pool=pool)
                segs.append(g)

                def cut(buckets):
                    segs[-1].capture_end()
                    plan.append(buckets)
                    nxt = torch.cuda.CUDAGraph()
                    nxt.capture_begin(pool=pool)
                    segs.append(nxt)

                def segmenter(ready):
                    at = {}
                    for b, (lo, hi, params) in enumerate(self._bucketed.spans):
                        idx = [ready[id(p)] for p in params if id(p) in ready]
                        at.setdefault(min(idx) if idx else -1, []).append(b)
                    return {i: (lambda bs=bs: cut(bs)) for i, bs in at.items()}
                self._segmenter = segmenter
                try:
                    out = self._fwd_bwd(st_sup, st_uns, ramp_val)
                finally:
                    self._segmenter = None
                segs[-1].capture_end()
            torch.cuda.current_stream().wait_stream(cap)
            plan.append([])
        else:
            g1 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1):
                out = self._fwd_bwd(st_sup, st_uns, ramp_val)
            segs, plan, pool = [g1], [None], g1.pool()
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g2, pool=pool):
            self._opt_ema()
        self.launches_per_replay = self.be.launches - l0
        self._restore(snap)
        self._graph = (segs, g2, out, float(ramp_val), tuple(tuple(t.shape) for t in sup_batch), plan)

    def _state_tensors(self):
        ts = list(self.student_net.state_dict().values())
        if self.teacher_net is not self.student_net:
            ts += list(self.teacher_net.state_dict().values())
        for net in (self.student_net, self.teacher_net):
            for m in net.modules():
                c = getattr(m, '_dev_counter', None)
                if c is not None:
                    ts.append(c)
        return ts

    def _snapshot(self):
        opt = {}
        for p, st in self.student_optim.state.items():
            opt[id(p)] = {k: v.clone() for k, v in st.items() if torch.is_tensor(v)}
        return [t.clone() for t in self._state_tensors()], opt

    def _restore(self, snap):
        """In-place restore (the captured graphs hold the addresses of these tensors)."""
        tensors, opt = snap
        with torch.no_grad():
            for dst, src in zip(self._state_tensors(), tensors):
                dst.copy_(src)
            for p, st in self.student_optim.state.items():
                old = opt.get(id(p))
                for k, v in st.items():
                    if torch.is_tensor(v):
                        if old is not None and k in old:
                            v.copy_(old[k])
                        else:
                            v.zero_()          # state created by the warm-up step: back to its initial value

    # -- input prefetch (graph mode): the NEXT call's pinned host batch is copied to staging buffers on a side stream while
    # this iteration computes; the next step() then only moves staging -> static buffers device to device.
    @staticmethod
    def _same_batch(a, b):
        (sa, ua), (sb, ub) = a, b
        if len(sa) != len(sb) or len(ua) != len(ub) or any(x is not y for x, y in zip(sa, sb)):
            return False
        return all(da.keys() == db.keys() and all(da[k] is db[k] for k in da) for da, db in zip(ua, ub))

    def _prefetch(self, sup_batch, unsup_batches):
        st_sup, st_uns = self._static
        if getattr(self, '_staging', None) is None:
            cache = {}

            def like(t):
                if id(t) not in cache:
                    cache[id(t)] = torch.empty_like(t)
                return cache[id(t)]
            self._staging = (tuple(like(t) for t in st_sup), [{k: like(v) for k, v in d.items()} for d in st_uns])
            self._side = torch.cuda.Stream()
            self._stage_free = None
        sg_sup, sg_uns = self._staging
        if len(sup_batch) != len(sg_sup) or len(unsup_batches) != len(sg_uns):
            return
        if self._stage_free is not None:
            self._side.wait_event(self._stage_free)          # the previous staging -> static copy has read the buffers
        with torch.cuda.stream(self._side):
            for dst, src in zip(sg_sup, sup_batch):
                dst.copy_(src, non_blocking=True)
            for d, ub in zip(sg_uns, unsup_batches):
                done = set()
                for k, v in ub.items():
                    if id(d[k]) not in done:
                        d[k].copy_(v, non_blocking=True)
                        done.add(id(d[k]))
            self._stage_done = torch.cuda.Event()
            self._stage_done.record(self._side)
        self._staged = (tuple(sup_batch), [dict(ub) for ub in unsup_batches])      # keeps the pinned tensors alive

    def _load_staged(self):
        st_sup, st_uns = self._static
        sg_sup, sg_uns = self._staging
        cur = torch.cuda.current_stream()
        cur.wait_event(self._stage_done)
        for dst, src in zip(st_sup, sg_sup):
            dst.copy_(src, non_blocking=True)
        for d, g in zip(st_uns, sg_uns):
            done = set()
            for k in d:
                if id(d[k]) not in done:
                    d[k].copy_(g[k], non_blocking=True)
                    done.add(id(d[k]))
        self._stage_free = torch.cuda.Event()
        self._stage_free.record(cur)
        self._staged = None

    def _load_static(self, sup_batch, unsup_batches):
        st_sup, st_uns = self._static
        for dst, src in zip(st_sup, sup_batch):
            dst.copy_(src, non_blocking=True)
        for d, ub in zip(st_uns, unsup_batches):
            done = set()
            for k, v in ub.items():
                if id(d[k]) not in done:
                    d[k].copy_(v, non_blocking=True)
                    done.add(id(d[k]))

    def _fwd_bwd(self, sup_batch, unsup_batches, ramp_val):
        from . import engine
        engine.invalidate_caches()       # derived weights (BN folds, dgrad transposes) are rebuilt once per iteration
        self._zero_grad()                                              # :290
        if self._can_batch_trunk(unsup_batches):
            return self._fwd_bwd_batched(sup_batch, unsup_batches[0], ramp_val)
        sup_loss = self.supervised(*sup_batch)
        cons, conf = None, None
        if self.cons_weight > 0.0:
            for ub in unsup_batches:
                if 'vat' in ub:
                    out4 = self.unsupervised_vat(ub['ux_tea'], ub['ux_stu'], ub['um'], ramp_val, ub.get('noise'))
                elif 'xf0_to_1' in ub:
                    out4 = self.unsupervised_aug(ub['ux0'], ub['um0'], ub['ux1'], ub['um1'], ub['xf0_to_1'], ramp_val)
                elif 'ict_mix_factors' in ub:
                    out4 = self.unsupervised_ict(ub['ux0_tea'], ub['ux0_stu'], ub['um0'], ub['ux1_tea'], ub['ux1_stu'],
                                                 ub['um1'], ub['ict_mix_factors'], ramp_val)
                elif self.mask_mix:
                    out4 = self.unsupervised_mix(ub['ux0_tea'], ub['ux0_stu'], ub['um0'], ub['ux1_tea'], ub['ux1_stu'],
                                                 ub['um1'], ub['mask_params'], ramp_val)
                else:
                    out4 = self.unsupervised_cut(ub['ux_tea'], ub['ux_stu'], ub['um'], ub['mask_params'], ramp_val)
                cons = out4[0] if cons is None else cons + out4[0]
                conf = out4[1] if conf is None else conf + out4[1]
        return {'sup_loss': sup_loss, 'cons_loss': cons, 'conf_rate': conf}

    def _opt_ema(self):
        if getattr(self.student_optim, 'b2_fused', False):
            self.student_optim.step(ema=self.teacher_optim)            # :465-467 in one launch
            return
        self.student_optim.step()                                      # :465
        if self.teacher_optim is not None:
            self.teacher_optim.step()                                  # :466-467

    def step(self, sup_batch, unsup_batches, ramp_val=1.0, eager=False, prefetch=None):
        """One full iteration.  `sup_batch` = (image, labels); `unsup_batches` = list (length
        unsup_batch_ratio) of dicts with keys ux0_tea, ux0_stu, um0, ux1_tea, ux1_stu, um1, mask_params (mix
        mode), the same with ict_mix_factors ((N,) fp32 Beta draws) instead of mask_params (ICT,
        train_seg_semisup_ict.py), or ux_tea, ux_stu, um, mask_params (cut mode), or ux0, um0, ux1, um1, xf0_to_1 ((N,2,3) fp32
        affine maps; augmentation consistency, train_seg_semisup_aug_mt.py), or ux_tea, ux_stu, um, vat (any value; VAT,
        train_seg_semisup_vat_mt.py; optional `noise` (N,3,H,W) replaces the torch.randn draw).  Returns device scalars
        {'sup_loss', 'cons_loss', 'conf_rate'} without synchronising.  With `use_cuda_graph` the iteration is captured
        once per (shapes, ramp value) and replayed; inputs may then be pinned host tensors (copied straight into the
        graph's static buffers).  `prefetch=(sup_batch, unsup_batches)`: the (pinned host) batch of the NEXT call; its
        host-to-device copy then runs on a side stream underneath this iteration (graph mode; ignored otherwise)."""
        if self.use_cuda_graph and not eager:
            key = (float(ramp_val), tuple(tuple(t.shape) for t in sup_batch))
            if self._graph is None or (self._graph[3], self._graph[4]) != key:
                self._capture(sup_batch, unsup_batches, ramp_val)
            staged = getattr(self, '_staged', None)
            if staged is not None and self._same_batch(staged, (sup_batch, unsup_batches)):
                self._load_staged()                    # prefetched during the previous call: device-to-device only
            else:
                self._load_static(sup_batch, unsup_batches)
            segs, g2, out, plan = self._graph[0], self._graph[1], self._graph[2], self._graph[5]
            for seg, buckets in zip(segs, plan):
                seg.replay()
                if buckets:                            # this segment completed these gradient buckets: all-reduce them under
                    for b in buckets:                  # the next segment
                        self._bucketed.launch(b)
                    self._overlapped = True
            self._allreduce()
            if hasattr(self.student_optim, 'upload_lr'):
                self.student_optim.upload_lr()         # this iteration's learning rates -> device, ahead of the captured step
            g2.replay()
            self.be.launches += self.launches_per_replay
            if prefetch is not None:
                self._prefetch(*prefetch)
            return out
        sup_batch = tuple(t.to(self._device(), non_blocking=True) for t in sup_batch)
        out = self._fwd_bwd(sup_batch, [{k: v.to(self._device(), non_blocking=True) for k, v in ub.items()} for ub in unsup_batches],
                            ramp_val)
        self._allreduce()
        self._opt_ema()
        return out

    def _device(self):
        return next(self.student_net.parameters()).device

"""ORACLE — test infrastructure only (see oracle/torch_oracle.py header).

Line-for-line CPU restatement of ONE iteration of the reference loop (train_seg_semisup_mask_mt.py:287-476)
on top of the functional networks of torch_oracle.py: construction (:86-134), iteration prologue (:287-290),
supervised branch (:296-301), mix (:309-369) / cut (:371-401) unsupervised branch, confidence + consistency
loss (:406-459), optimiser and EMA steps (:465-467).  Pinned against the real reference modules by
oracle/gen_golden.py (the reference's job function itself cannot run offline: datapipe needs skimage and
real datasets).
"""
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch_oracle as TO  # noqa: E402


def _is_bn_key(k):
    return k.endswith(('.running_mean', '.running_var', '.num_batches_tracked'))


def deeplab2_trainable(key):
    """requires_grad of a DeepLab v2 state entry: conv weights / layer5 biases only (BN affine frozen,
    deeplab2.py:72-84, 141-143, 167-168)."""
    if _is_bn_key(key):
        return False
    if 'bn' in key.split('.')[-2] or '.downsample.1.' in key:
        return False
    return True


def deeplab2_param_groups(sd):
    """[pretrained (0.1 lr) with the reference's repetitions, new] — deeplab2.py:208-242.  A tensor under
    `top` is yielded once per enclosing module of `top`'s module tree (modules() x parameters())."""
    g0 = []
    for k, v in sd.items():
        if not deeplab2_trainable(k) or k.startswith('layer5.'):
            continue
        parts = k.split('.')[:-1]             # module path of the leaf module
        g0.extend([v] * len(parts))           # conv1 -> 1, layerX.i.convY -> 3, layerX.i.downsample.0 -> 4
    g1 = [v for k, v in sd.items() if k.startswith('layer5.') and not _is_bn_key(k)]
    return g0, g1


class OracleMeanTeacher(object):
    """State + one-iteration function.  arch: 'deeplab2' | 'deeplab3plus' | 'deeplab3' | 'resunet' | 'denseunet'."""

    def __init__(self, arch, state_dict, learning_rate, opt_type='adam', teacher_alpha=0.99, freeze_bn=True,
                 cons_loss_fn='var', cons_weight=1.0, conf_thresh=0.97, conf_per_pixel=False, rampup=-1, mask_mix=True,
                 dtype=torch.float32, model='mean_teacher', vat_radius=0.5, adaptive_vat_radius=False,
                 vat_dir_from_student=False):
        self.arch, self.freeze_bn = arch, freeze_bn
        self.cons_loss_fn, self.cons_weight = cons_loss_fn, cons_weight
        self.conf_thresh, self.conf_per_pixel, self.rampup, self.mask_mix = conf_thresh, conf_per_pixel, rampup, mask_mix
        self.teacher_alpha = teacher_alpha
        self.vat_radius, self.adaptive_vat_radius, self.vat_dir_from_student = vat_radius, adaptive_vat_radius, vat_dir_from_student
        # nn.Module.training of the two networks; VAT's direction pass switches one of them to eval mode and the reference
        # never switches it back inside the epoch (train_seg_semisup_vat_mt.py:237 vs :326-333)
        self.eval_mode = {'student': False, 'teacher': False}

        def clone(sd):
            return OrderedDict((k, (v.to(dtype) if v.dtype == torch.float32 else v).clone()) for k, v in sd.items())
        self.student = clone(state_dict)
        self.model = model
        if model == 'pi':                         # :110-113: teacher_net = student_net, no teacher optimiser
            self.teacher = self.student
        else:
            self.teacher = clone(state_dict)      # EMAWeightOptimizer.__init__ copies student -> teacher (:12-13)
        if arch == 'deeplab2':
            for k, v in self.student.items():
                if deeplab2_trainable(k):
                    v.requires_grad_(True)
            g0, g1 = deeplab2_param_groups(self.student)
        else:
            for k, v in self.student.items():
                if not _is_bn_key(k):
                    v.requires_grad_(True)
            g0, g1 = [], [v for k, v in self.student.items() if not _is_bn_key(k)]   # deeplab3plus.py:138-151
        groups = [dict(params=g0, lr=learning_rate * 0.1), dict(params=g1, lr=learning_rate)]   # :90-93
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            if opt_type == 'adam':
                self.optim = torch.optim.Adam(groups, foreach=False)
            else:
                self.optim = torch.optim.SGD(groups, momentum=0.9, nesterov=False, weight_decay=5e-4, foreach=False)

    # ------------------------------------------------------------------------------------------
    def start_epoch(self):
        """student_net.train(); teacher_net.train() (+ freeze_batchnorm) at the top of every epoch (:326-333)."""
        self.eval_mode = {'student': False, 'teacher': False}

    def _forward(self, sd, x, dropout_masks=None):
        which = 'student' if sd is self.student else 'teacher'
        if self.eval_mode[which]:                 # module in eval mode: running statistics everywhere, no dropout
            if self.arch == 'deeplab2':
                return TO.deeplab2_forward(sd, x, bn_train=False)
            if self.arch == 'denseunet':
                return TO.denseunet_forward(sd, x, backbone_bn_train=False, head_bn_train=False)
            if self.arch == 'resunet':
                return TO.resunet_forward(sd, x, backbone_bn_train=False, head_bn_train=False)
            fwd = TO.deeplab3_forward if self.arch == 'deeplab3' else TO.deeplab3plus_forward
            return fwd(sd, x, backbone_bn_train=False, head_bn_train=False)
        if self.arch == 'deeplab2':
            return TO.deeplab2_forward(sd, x, bn_train=not self.freeze_bn)
        if self.arch == 'denseunet':              # architectures/denseunet.py: freeze_batchnorm touches the encoder only
            return TO.denseunet_forward(sd, x, backbone_bn_train=not self.freeze_bn, head_bn_train=True,
                                        dropout_masks=dropout_masks)
        if self.arch == 'resunet':                # architectures/resunet.py: freeze_batchnorm touches the encoder only
            return TO.resunet_forward(sd, x, backbone_bn_train=not self.freeze_bn, head_bn_train=True,
                                      dropout_masks=dropout_masks)
        if self.arch == 'deeplab3':               # torchvision DeepLab v3 in the reference's wrapper (same BN-freezing scope)
            return TO.deeplab3_forward(sd, x, backbone_bn_train=not self.freeze_bn, head_bn_train=True,
                                       dropout_masks=dropout_masks)
        # net.train() then freeze_batchnorm(): backbone BN eval, head BN train (:268-275, deeplab3plus.py:120-121)
        return TO.deeplab3plus_forward(sd, x, backbone_bn_train=not self.freeze_bn, head_bn_train=True,
                                       dropout_masks=dropout_masks)

    def ema_step(self):
        """optim_weight_ema.py:21-25 over every float state tensor (parameters and BN running statistics)."""
        if self.model == 'pi':                    # :466 `if teacher_optim is not None`
            return
        a = self.teacher_alpha
        one_minus = 1.0 - a
        with torch.no_grad():
            for k, t in self.teacher.items():
                if t.dtype.is_floating_point:
                    t.mul_(a)
                    t.add_(self.student[k].detach() * one_minus)

    def step(self, sup_x, sup_y, unsup, ramp_val=1.0, drop=None):
        """unsup: dict like cutmix_semisup_seg_b200.synthetic.make_unsup_batch with DENSE mask params
        (N,1,H,W).  drop: optional dict of explicit dropout keep-masks (N,256,h,w) for the four forward passes:
        {'sup','tea0','tea1','stu'}.  Returns python floats (sup_loss, cons_loss, conf_rate)."""
        drop = drop or {}
        self.optim.zero_grad()                                                      # :290
        logits_sup = self._forward(self.student, sup_x, drop.get('sup'))            # :299
        sup_loss = TO.supervised_loss(logits_sup, sup_y)                            # :300
        sup_loss.backward()                                                         # :301
        cons_val, conf_val = 0.0, 0.0
        unsup_list = unsup if isinstance(unsup, (list, tuple)) else [unsup]        # :304 `for _ in range(unsup_batch_ratio)`
        for unsup in (unsup_list if self.cons_weight > 0.0 else []):
            m = unsup.get('mask_params')
            if 'vat' in unsup:                                                      # train_seg_semisup_vat_mt.py:364-452
                which = 'student' if (self.vat_dir_from_student or self.model == 'pi') else 'teacher'    # :102-105
                dir_sd = self.student if which == 'student' else self.teacher
                self.eval_mode[which] = True                                        # vat_dir_net.eval() :237 (persists)
                x_perturb = TO.vat_perturbation(lambda t: self._forward(dir_sd, t), unsup['ux_tea'], unsup['ux_stu'],
                                                unsup['noise'], self.cons_loss_fn, self.vat_radius, self.adaptive_vat_radius)
                ux_adv = unsup['ux_stu'] + x_perturb                                # :392
                with torch.no_grad():                                               # :395-396
                    lt = self._forward(self.teacher, unsup['ux_tea'], drop.get('tea0')).detach()
                ls = self._forward(self.student, ux_adv, drop.get('stu'))           # :398
                loss, conf = TO.consistency_loss(lt, None, ls, None, unsup['um'], self.cons_loss_fn, self.conf_thresh,
                                                 self.conf_per_pixel, ramp_val, self.rampup)
            elif 'xf0_to_1' in unsup:                                               # train_seg_semisup_aug_mt.py:275-398
                with torch.no_grad():                                               # aug :291-293
                    lt = self._forward(self.teacher, unsup['ux0'], drop.get('tea0')).detach()
                ls = self._forward(self.student, unsup['ux1'], drop.get('stu'))     # aug :295
                loss, conf = TO.aug_consistency_loss(lt, ls, unsup['xf0_to_1'], unsup['um0'], unsup['um1'], self.cons_loss_fn,
                                                     self.conf_thresh, self.conf_per_pixel, ramp_val, self.rampup)
            elif 'ict_mix_factors' in unsup:                                        # train_seg_semisup_ict.py:306-392
                f = unsup['ict_mix_factors'].reshape(-1, 1, 1, 1)
                ux_mixed = unsup['ux0_stu'] * (1.0 - f) + unsup['ux1_stu'] * f      # ict :310
                um_mixed = unsup['um0'] * (1.0 - f) + unsup['um1'] * f              # ict :311
                with torch.no_grad():                                               # ict :314-316
                    l0 = self._forward(self.teacher, unsup['ux0_tea'], drop.get('tea0')).detach()
                    l1 = self._forward(self.teacher, unsup['ux1_tea'], drop.get('tea1')).detach()
                ls = self._forward(self.student, ux_mixed, drop.get('stu'))         # ict :318
                loss, conf = TO.ict_consistency_loss(l0, l1, ls, f, um_mixed, self.cons_loss_fn, self.conf_thresh,
                                                     self.conf_per_pixel, ramp_val, self.rampup)
            elif self.mask_mix:
                ux_mixed = unsup['ux0_stu'] * (1 - m) + unsup['ux1_stu'] * m        # :350
                um_mixed = unsup['um0'] * (1 - m) + unsup['um1'] * m                # :351
                with torch.no_grad():                                               # :354-356
                    l0 = self._forward(self.teacher, unsup['ux0_tea'], drop.get('tea0')).detach()
                    l1 = self._forward(self.teacher, unsup['ux1_tea'], drop.get('tea1')).detach()
                ls = self._forward(self.student, ux_mixed, drop.get('stu'))         # :358
                loss, conf = TO.consistency_loss(l0, l1, ls, m, um_mixed, self.cons_loss_fn, self.conf_thresh,
                                                 self.conf_per_pixel, ramp_val, self.rampup)
            else:
                ux_cut = unsup['ux_stu'] * m                                        # :389
                with torch.no_grad():
                    lt = self._forward(self.teacher, unsup['ux_tea'], drop.get('tea0')).detach()
                ls = self._forward(self.student, ux_cut, drop.get('stu'))
                loss, conf = TO.consistency_loss(lt, None, ls, None, m * unsup['um'], self.cons_loss_fn, self.conf_thresh,
                                                 self.conf_per_pixel, ramp_val, self.rampup)
            (loss * self.cons_weight).backward()                                    # :458-459
            cons_val += float(loss.detach()); conf_val += float(conf)               # :461 / :413 accumulators
        self.optim.step()                                                           # :465
        self.ema_step()                                                             # :466-467
        return float(sup_loss.detach()), cons_val, conf_val

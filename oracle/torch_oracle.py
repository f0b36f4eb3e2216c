"""ORACLE — test infrastructure only (never imported by the product path).

A plain PyTorch fp32 (or fp64) CPU restatement of the reference's CutMix mean-teacher hot path, written
functionally on top of a `state_dict`, so that it travels to the GPU box where /root/reference does
not exist.  Each function cites the reference lines it follows.  It is pinned against the real reference
(imported from /root/reference in the development container) by oracle/gen_golden.py, whose outputs are
committed under tests/golden/ and re-checked by tests/test_oracle_golden.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# Building blocks
# ----------------------------------------------------------------------------------------------
def _bn(sd, prefix, x, train, momentum=0.1, eps=1e-5):
    """nn.BatchNorm2d forward; `train` selects batch statistics (and updates the running buffers in
    place, like the module does)."""
    return F.batch_norm(x, sd[prefix + '.running_mean'], sd[prefix + '.running_var'], sd[prefix + '.weight'],
                        sd[prefix + '.bias'], training=train, momentum=momentum, eps=eps)


def _count_bn(sd, prefix, train):
    if train:
        sd[prefix + '.num_batches_tracked'] += 1


# ----------------------------------------------------------------------------------------------
# DeepLab v2 (reference architectures/deeplab2.py)
# ----------------------------------------------------------------------------------------------
_DL2_LAYERS = (('layer1', 64, 3, 1, 1), ('layer2', 128, 4, 2, 1), ('layer3', 256, 23, 1, 2), ('layer4', 512, 3, 1, 4))


def deeplab2_forward(sd, x, bn_train=False):
    """ResNetDeepLab.forward (deeplab2.py:183-206): stem, ceil-mode max-pool (:146), bottlenecks with the
    stride on conv1 (:70) and dilated conv2 (:76-77), layer5 = conv(d=6) + conv(d=12) only (:124-128),
    bilinear align_corners=True up-sampling to the input size (:204)."""
    in_hw = x.shape[2:4]
    t = F.conv2d(x, sd['conv1.weight'], stride=2, padding=3)
    t = F.relu(_bn(sd, 'bn1', t, bn_train)); _count_bn(sd, 'bn1', bn_train)
    t = F.max_pool2d(t, 3, 2, 1, ceil_mode=True)
    for name, planes, blocks, stride, dil in _DL2_LAYERS:
        b = -1
        while '{}.{}.conv1.weight'.format(name, b + 1) in sd:        # block count from the keys (ResNet-101: 3,4,23,3)
            b += 1
            p = '{}.{}'.format(name, b)
            s = stride if b == 0 else 1
            res = t
            o = F.conv2d(t, sd[p + '.conv1.weight'], stride=s)
            o = F.relu(_bn(sd, p + '.bn1', o, bn_train)); _count_bn(sd, p + '.bn1', bn_train)
            o = F.conv2d(o, sd[p + '.conv2.weight'], padding=dil, dilation=dil)
            o = F.relu(_bn(sd, p + '.bn2', o, bn_train)); _count_bn(sd, p + '.bn2', bn_train)
            o = F.conv2d(o, sd[p + '.conv3.weight'])
            o = _bn(sd, p + '.bn3', o, bn_train); _count_bn(sd, p + '.bn3', bn_train)
            if (p + '.downsample.0.weight') in sd:
                res = F.conv2d(t, sd[p + '.downsample.0.weight'], stride=s)
                res = _bn(sd, p + '.downsample.1', res, bn_train); _count_bn(sd, p + '.downsample.1', bn_train)
            t = F.relu(o + res)
    out = F.conv2d(t, sd['layer5.conv2d_list.0.weight'], sd['layer5.conv2d_list.0.bias'], padding=6, dilation=6)
    out = out + F.conv2d(t, sd['layer5.conv2d_list.1.weight'], sd['layer5.conv2d_list.1.bias'], padding=12, dilation=12)
    return F.interpolate(out, size=in_hw, mode='bilinear', align_corners=True)


# ----------------------------------------------------------------------------------------------
# DeepLab v3+ (reference architectures/deeplab3plus.py + torchvision ResNet-101 / ASPP)
# ----------------------------------------------------------------------------------------------
# torchvision resnet101(replace_stride_with_dilation=[False, True, True]): (name, planes, blocks, stride, first
# block dilation, remaining blocks dilation)
_DL3_LAYERS = (('layer1', 64, 3, 1, 1, 1), ('layer2', 128, 4, 2, 1, 1), ('layer3', 256, 23, 1, 1, 2),
               ('layer4', 512, 3, 1, 2, 4))


def deeplab3_forward(sd, x, backbone_bn_train=False, head_bn_train=False, dropout_masks=None, pre='deeplab.'):
    """DeepLabv3Wrapper.forward on torchvision's DeepLab v3 (`deeplabv3_resnet101`, reference network_architectures.py:75-98):
    the same backbone and ASPP as DeepLab v3+, then DeepLabHead's 3x3 conv -> BN -> ReLU -> 1x1 conv and one bilinear resize."""
    return deeplab3plus_forward(sd, x, backbone_bn_train, head_bn_train, dropout_masks, pre, head='v3')


def deeplab3plus_forward(sd, x, backbone_bn_train=False, head_bn_train=False, dropout_masks=None, pre='deeplab.',
                         aspp_rates=(12, 24, 36), head='v3plus'):
    """DeepLabv3Wrapper.forward -> DeepLabV3Plus.forward -> DeepLabHeadV3Plus.forward
    (deeplab3plus.py:116-117, 73-78, 51-56) with torchvision's ASPP.  `dropout_masks`: list with one
    (N,256,h,w) keep-mask for the ASPP Dropout(0.5) (None = dropout inactive)."""
    in_hw = x.shape[2:4]
    bb = pre + 'backbone.'
    t = F.conv2d(x, sd[bb + 'conv1.weight'], stride=2, padding=3)
    t = F.relu(_bn(sd, bb + 'bn1', t, backbone_bn_train)); _count_bn(sd, bb + 'bn1', backbone_bn_train)
    t = F.max_pool2d(t, 3, 2, 1)
    feats = {}
    for name, planes, blocks, stride, d0, d1 in _DL3_LAYERS:
        b = -1
        while '{}{}.{}.conv1.weight'.format(bb, name, b + 1) in sd:
            b += 1
            p = '{}{}.{}'.format(bb, name, b)
            s = stride if b == 0 else 1
            dil = d0 if b == 0 else d1
            res = t
            o = F.conv2d(t, sd[p + '.conv1.weight'])
            o = F.relu(_bn(sd, p + '.bn1', o, backbone_bn_train)); _count_bn(sd, p + '.bn1', backbone_bn_train)
            o = F.conv2d(o, sd[p + '.conv2.weight'], stride=s, padding=dil, dilation=dil)
            o = F.relu(_bn(sd, p + '.bn2', o, backbone_bn_train)); _count_bn(sd, p + '.bn2', backbone_bn_train)
            o = F.conv2d(o, sd[p + '.conv3.weight'])
            o = _bn(sd, p + '.bn3', o, backbone_bn_train); _count_bn(sd, p + '.bn3', backbone_bn_train)
            if (p + '.downsample.0.weight') in sd:
                res = F.conv2d(t, sd[p + '.downsample.0.weight'], stride=s)
                res = _bn(sd, p + '.downsample.1', res, backbone_bn_train); _count_bn(sd, p + '.downsample.1', backbone_bn_train)
            t = F.relu(o + res)
        feats[name] = t
    hd = pre + 'classifier.'
    tr = head_bn_train

    def cbr(x_, cp, bp, **kw):
        y = F.conv2d(x_, sd[cp + '.weight'], **kw)
        y = F.relu(_bn(sd, bp, y, tr)); _count_bn(sd, bp, tr)
        return y
    ap = hd + ('aspp.' if head == 'v3plus' else '0.')          # DeepLabHeadV3Plus.aspp / DeepLabHead[0]
    if head == 'v3plus':
        low = cbr(feats['layer1'], hd + 'project.0', hd + 'project.1')
    f = feats['layer4']
    branches = [cbr(f, ap + 'convs.0.0', ap + 'convs.0.1')]
    n_branches = 1 + len(aspp_rates)
    for i in range(1, n_branches):
        rate = aspp_rates[i - 1]
        branches.append(cbr(f, ap + 'convs.{}.0'.format(i), ap + 'convs.{}.1'.format(i), padding=rate, dilation=rate))
    pi = n_branches
    pooled = F.adaptive_avg_pool2d(f, 1)
    pooled = cbr(pooled, ap + 'convs.{}.1'.format(pi), ap + 'convs.{}.2'.format(pi))
    branches.append(F.interpolate(pooled, size=f.shape[2:4], mode='bilinear', align_corners=False))
    a = cbr(torch.cat(branches, dim=1), ap + 'project.0', ap + 'project.1')
    if dropout_masks is not None:
        a = a * dropout_masks[0] * 2.0            # nn.Dropout(0.5) in training mode
    if head == 'v3':                              # torchvision DeepLabHead[1:] + DeepLabV3.forward's resize
        c = cbr(a, hd + '1', hd + '2', padding=1)
        c = F.conv2d(c, sd[hd + '4.weight'], sd[hd + '4.bias'])
        return F.interpolate(c, size=in_hw, mode='bilinear', align_corners=False)
    a = F.interpolate(a, size=low.shape[2:4], mode='bilinear', align_corners=False)
    c = torch.cat([low, a], dim=1)
    c = cbr(c, hd + 'classifier.0', hd + 'classifier.1', padding=1)
    c = cbr(c, hd + 'classifier.3', hd + 'classifier.4', padding=1)
    c = F.conv2d(c, sd[hd + 'classifier.6.weight'], sd[hd + 'classifier.6.bias'])
    return F.interpolate(c, size=in_hw, mode='bilinear', align_corners=False)


def resunet_forward(sd, x, backbone_bn_train=False, head_bn_train=True, dropout_masks=None):
    """ResUNet.forward, reference architectures/resunet.py:66-92, on a torchvision ResNet-50 / -101 state_dict under
    `base_model.` (stride on the 3x3 convolution of each bottleneck, no dilation).  `dropout_masks`: list with one (N,64,H,W)
    keep-mask for `final_dec_drop` = nn.Dropout(0.3) (None = dropout inactive)."""
    bb = 'base_model.'
    t = F.conv2d(x, sd[bb + 'conv1.weight'], stride=2, padding=3)                                 # :67
    t = _bn(sd, bb + 'bn1', t, backbone_bn_train); _count_bn(sd, bb + 'bn1', backbone_bn_train)   # :68 `r2 = x = bn1(x)`
    r2 = t = F.relu(t)               # :69 base_model.relu is nn.ReLU(inplace=True): it rectifies the tapped tensor r2 too
    t = F.max_pool2d(t, 3, 2, 1)                                                                  # :70
    taps = []
    for name, stride in (('layer1', 1), ('layer2', 2), ('layer3', 2), ('layer4', 2)):             # :72-75
        b = -1
        while '{}{}.{}.conv1.weight'.format(bb, name, b + 1) in sd:
            b += 1
            p = '{}{}.{}'.format(bb, name, b)
            s_ = stride if b == 0 else 1
            res = t
            o = F.conv2d(t, sd[p + '.conv1.weight'])
            o = F.relu(_bn(sd, p + '.bn1', o, backbone_bn_train)); _count_bn(sd, p + '.bn1', backbone_bn_train)
            o = F.conv2d(o, sd[p + '.conv2.weight'], stride=s_, padding=1)
            o = F.relu(_bn(sd, p + '.bn2', o, backbone_bn_train)); _count_bn(sd, p + '.bn2', backbone_bn_train)
            o = F.conv2d(o, sd[p + '.conv3.weight'])
            o = _bn(sd, p + '.bn3', o, backbone_bn_train); _count_bn(sd, p + '.bn3', backbone_bn_train)
            if (p + '.downsample.0.weight') in sd:
                res = F.conv2d(t, sd[p + '.downsample.0.weight'], stride=s_)
                res = _bn(sd, p + '.downsample.1', res, backbone_bn_train); _count_bn(sd, p + '.downsample.1', backbone_bn_train)
            t = F.relu(o + res)
        taps.append(t)
    r4, r8, r16, r32 = taps
    t = F.conv2d(r32, sd['line0_conv.weight'], sd['line0_conv.bias'])                             # :78
    for name, skip in (('decoder3', r16), ('decoder2', r8), ('decoder1', r4), ('decoder0', r2)):  # :81-84, :27-33
        t = F.interpolate(t, scale_factor=2, mode='nearest') + skip
        t = F.conv2d(t, sd[name + '.conv.weight'], padding=1)
        t = F.relu(_bn(sd, name + '.conv_bn', t, head_bn_train)); _count_bn(sd, name + '.conv_bn', head_bn_train)
    t = F.conv2d(F.interpolate(t, scale_factor=2, mode='nearest'), sd['final_dec_conv.weight'], padding=1)   # :87
    if dropout_masks is not None:
        t = t * dropout_masks[0] * (1.0 / (1.0 - 0.3))                                            # nn.Dropout(0.3), training
    t = F.relu(_bn(sd, 'final_dec_bn', t, head_bn_train)); _count_bn(sd, 'final_dec_bn', head_bn_train)      # :87-88
    return F.conv2d(t, sd['final_clf.weight'], sd['final_clf.bias'])                              # :89


def denseunet_forward(sd, x, backbone_bn_train=False, head_bn_train=True, dropout_masks=None):
    """DenseUNet.forward, reference architectures/denseunet.py:98-124, on a torchvision DenseNet-161 state_dict under
    `base_model.features.` (pre-activation dense layers: norm1-relu-conv1(1x1)-norm2-relu-conv2(3x3), concatenated;
    transitions norm-relu-conv(1x1)-avgpool(2)).  Taps are taken BEFORE pool0 / transition1..3 (:100-104).
    `dropout_masks`: one (N,64,H,W) keep-mask for `final_dec_drop` = nn.Dropout(0.3) (None = inactive)."""
    ft = 'base_model.features.'
    tr = backbone_bn_train

    def bn_relu(t, prefix):
        y = F.relu(_bn(sd, prefix, t, tr)); _count_bn(sd, prefix, tr)
        return y
    t = F.conv2d(x, sd[ft + 'conv0.weight'], stride=2, padding=3)
    t = bn_relu(t, ft + 'norm0')                       # norm0, relu0 (in place); the 'pool0' tap is this tensor
    enc_x = [t]
    t = F.max_pool2d(t, 3, 2, 1)
    for b in (1, 2, 3, 4):
        feats = [t]
        i = 1
        while '{}denseblock{}.denselayer{}.conv1.weight'.format(ft, b, i) in sd:
            p = '{}denseblock{}.denselayer{}.'.format(ft, b, i)
            o = bn_relu(torch.cat(feats, 1), p + 'norm1')
            o = F.conv2d(o, sd[p + 'conv1.weight'])
            o = bn_relu(o, p + 'norm2')
            feats.append(F.conv2d(o, sd[p + 'conv2.weight'], padding=1))
            i += 1
        t = torch.cat(feats, 1)
        if b < 4:
            enc_x.append(t)                            # tap before transition b
            p = '{}transition{}.'.format(ft, b)
            t = F.avg_pool2d(F.conv2d(bn_relu(t, p + 'norm'), sd[p + 'conv.weight']), 2, 2)
    t = _bn(sd, ft + 'norm5', t, tr); _count_bn(sd, ft + 'norm5', tr)
    t = F.relu(t)                                                                                 # :108
    enc_x[-1] = F.conv2d(enc_x[-1], sd['line0_conv.weight'], sd['line0_conv.bias'])               # :111-112
    for k, skip in zip((3, 2, 1, 0), enc_x[::-1]):                                                # :115-116 (stored reversed, :95)
        name = 'decoder_blocks.{}'.format(k)
        t = F.interpolate(t, scale_factor=2, mode='nearest') + skip
        t = F.conv2d(t, sd[name + '.conv.weight'], padding=1)
        t = F.relu(_bn(sd, name + '.conv_bn', t, head_bn_train)); _count_bn(sd, name + '.conv_bn', head_bn_train)
    t = F.conv2d(F.interpolate(t, scale_factor=2, mode='nearest'), sd['final_dec_conv.weight'], padding=1)   # :119
    if dropout_masks is not None:
        t = t * dropout_masks[0] * (1.0 / (1.0 - 0.3))
    t = F.relu(_bn(sd, 'final_dec_bn', t, head_bn_train)); _count_bn(sd, 'final_dec_bn', head_bn_train)      # :119-120
    return F.conv2d(t, sd['final_clf.weight'], sd['final_clf.bias'])                              # :121


# ----------------------------------------------------------------------------------------------
# Loss block (reference train_seg_semisup_mask_mt.py:363-367, 406-459) and CE (:126, :300)
# ----------------------------------------------------------------------------------------------
def consistency_loss(logits_tea0, logits_tea1, logits_stu, mix_mask, loss_mask, cons_loss_fn='var', conf_thresh=0.97,
                     conf_per_pixel=False, ramp_val=1.0, rampup=-1):
    """Returns (consistency_loss [what the reference logs, :461], conf_rate, loss tensor to back-propagate
    before multiplication by cons_weight)."""
    if logits_tea1 is not None:
        logits_cons_tea = logits_tea0 * (1 - mix_mask) + logits_tea1 * mix_mask       # :363
    else:
        logits_cons_tea = logits_tea0
    prob_tea = F.softmax(logits_cons_tea, dim=1)                                       # :366
    prob_stu = F.softmax(logits_stu, dim=1)                                            # :367
    n_classes = logits_stu.shape[1]
    conf_rate = torch.tensor(float('nan'))
    if conf_thresh > 0.0:                                                              # :407-418
        conf_tea = prob_tea.max(dim=1)[0]
        conf_mask = (conf_tea >= conf_thresh).float()[:, None, :, :]
        conf_rate = conf_mask.mean()
        if not conf_per_pixel:
            conf_mask = conf_mask.mean()
        loss_mask = loss_mask * conf_mask
    if cons_loss_fn == 'var':                                                          # :428-431
        d = prob_stu - prob_tea
        q = (d * d).sum(dim=1, keepdim=True)
    elif cons_loss_fn == 'logits_var':                                                 # :432-435
        d = logits_stu - logits_cons_tea
        q = (d * d).sum(dim=1, keepdim=True) / math.sqrt(n_classes)
    elif cons_loss_fn == 'logits_smoothl1':                                            # :436-439
        q = F.smooth_l1_loss(logits_stu, logits_cons_tea, reduction='none').sum(dim=1, keepdim=True) / math.sqrt(n_classes)
    elif cons_loss_fn == 'bce':                                                        # :440-443
        eps = 1e-6
        q = -(prob_tea * torch.log(prob_stu + eps) + (1.0 - prob_tea) * torch.log(1.0 - prob_stu + eps))
        q = q.sum(dim=1, keepdim=True)
    elif cons_loss_fn == 'kld':                                                        # :444-446
        q = F.kl_div(F.log_softmax(logits_stu, dim=1), prob_tea, reduction='none').sum(dim=1, keepdim=True)
    else:
        raise ValueError(cons_loss_fn)
    loss = (q * loss_mask).mean()                                                      # :451
    if rampup > 0:
        loss = loss * ramp_val                                                         # :454-455
    return loss, conf_rate


def ict_consistency_loss(logits_u0_tea, logits_u1_tea, logits_cons_stu, ict_mix_factors, loss_mask, cons_loss_fn='var',
                         conf_thresh=0.97, conf_per_pixel=False, ramp_val=1.0, rampup=-1):
    """ICT loss block, reference train_seg_semisup_ict.py:320-386 line by line.  `ict_mix_factors`: (N,1,1,1) float32;
    `loss_mask` = the mixed valid mask (:332).  Returns (consistency_loss, conf_rate) like consistency_loss().

    Kept on purpose: `conf_mask[:, None, :, :]` (:344) is applied to an (N,1,H,W) tensor (max(..., keepdim=True), :338-339),
    so with --conf_per_pixel the product with the (N,1,H,W) loss mask broadcasts to (N,N,1,H,W)."""
    prob_u0_tea = F.softmax(logits_u0_tea, dim=1)                                      # :320
    prob_u1_tea = F.softmax(logits_u1_tea, dim=1)                                      # :321
    prob_cons_stu = F.softmax(logits_cons_stu, dim=1)                                  # :322
    logits_cons_tea = logits_u0_tea * (1 - ict_mix_factors) + logits_u1_tea * ict_mix_factors   # :328
    prob_cons_tea = prob_u0_tea * (1 - ict_mix_factors) + prob_u1_tea * ict_mix_factors         # :329
    n_classes = logits_cons_stu.shape[1]
    conf_rate = torch.tensor(float('nan'))
    if conf_thresh > 0.0:                                                              # :335-351
        conf_u0_tea = prob_u0_tea.max(dim=1, keepdim=True)[0]
        conf_u1_tea = prob_u1_tea.max(dim=1, keepdim=True)[0]
        conf_tea = conf_u0_tea * (1 - ict_mix_factors) + conf_u1_tea * ict_mix_factors
        conf_mask = (conf_tea >= conf_thresh).float()[:, None, :, :]
        conf_rate = conf_mask.mean()
        if not conf_per_pixel:
            conf_mask = conf_mask.mean()
        loss_mask = loss_mask * conf_mask
    if cons_loss_fn == 'var':                                                          # :360-363
        d = prob_cons_stu - prob_cons_tea
        q = (d * d).sum(dim=1, keepdim=True)
    elif cons_loss_fn == 'logits_var':                                                 # :364-367
        d = logits_cons_stu - logits_cons_tea
        q = (d * d).sum(dim=1, keepdim=True) / math.sqrt(n_classes)
    elif cons_loss_fn == 'logits_smoothl1':                                            # :368-371
        q = F.smooth_l1_loss(logits_cons_stu, logits_cons_tea, reduction='none').sum(dim=1, keepdim=True) / math.sqrt(n_classes)
    elif cons_loss_fn == 'bce':                                                        # :372-375
        eps = 1e-6
        q = -(prob_cons_tea * torch.log(prob_cons_stu + eps) + (1.0 - prob_cons_tea) * torch.log(1.0 - prob_cons_stu + eps))
        q = q.sum(dim=1, keepdim=True)
    elif cons_loss_fn == 'kld':                                                        # :376-378
        q = F.kl_div(F.log_softmax(logits_cons_stu, dim=1), prob_cons_tea, reduction='none').sum(dim=1, keepdim=True)
    else:
        raise ValueError(cons_loss_fn)
    loss = (q * loss_mask).mean()                                                      # :383
    if rampup > 0:
        loss = loss * ramp_val                                                         # :386-387
    return loss, conf_rate


def aug_consistency_loss(logits_cons_tea, logits_cons_stu, xf0_to_1, um0, um1, cons_loss_fn='var', conf_thresh=0.97,
                         conf_per_pixel=False, ramp_val=1.0, rampup=-1, strict_reference=True):
    """Augmentation-driven consistency block, reference train_seg_semisup_aug_mt.py:302-391 line by line: the teacher's
    logits, soft-max probabilities and valid mask are resampled into the student's frame with
    F.affine_grid / F.grid_sample (align_corners=True: datapipe/torch_utils.py:10-12 on torch >= 1.3).  `xf0_to_1`: (N,2,3).
    Returns (consistency_loss, conf_rate) like consistency_loss().

    Kept on purpose: the reference's `logits_var` branch (:370-374) overwrites its result with `delta_prob * delta_prob`,
    a variable only the `var` branch assigns, so it raises UnboundLocalError; `strict_reference=False` evaluates the
    formula of the sibling scripts instead (used to check the kernel's loss_fn = 1 path)."""
    kw = dict(align_corners=True)
    grid_tea_to_stu = F.affine_grid(xf0_to_1, logits_cons_tea.shape[:1] + (3,) + logits_cons_tea.shape[2:], **kw)   # :302
    logits_cons_tea_in_stu = F.grid_sample(logits_cons_tea, grid_tea_to_stu, **kw)                  # :304
    mask_tea_in_stu = F.grid_sample(um0, grid_tea_to_stu, **kw) * um1                               # :306
    prob_cons_tea = F.softmax(logits_cons_tea, dim=1)                                               # :309
    prob_cons_stu = F.softmax(logits_cons_stu, dim=1)                                               # :310
    prob_cons_tea_in_stu = F.grid_sample(prob_cons_tea, grid_tea_to_stu, **kw)                      # :312
    loss_mask = mask_tea_in_stu                                                                     # :341
    n_classes = logits_cons_stu.shape[1]
    conf_rate = torch.tensor(float('nan'))
    if conf_thresh > 0.0:                                                                           # :345-356
        conf_tea = prob_cons_tea_in_stu.max(dim=1)[0]
        conf_mask = (conf_tea >= conf_thresh).float()[:, None, :, :]
        conf_rate = conf_mask.mean()
        if not conf_per_pixel:
            conf_mask = conf_mask.mean()
        loss_mask = loss_mask * conf_mask
    if cons_loss_fn == 'var':                                                                       # :366-369
        d = prob_cons_stu - prob_cons_tea_in_stu
        q = (d * d).sum(dim=1, keepdim=True)
    elif cons_loss_fn == 'logits_var':                                                              # :370-374
        if strict_reference:
            raise UnboundLocalError("local variable 'delta_prob' referenced before assignment")
        d = logits_cons_stu - logits_cons_tea_in_stu
        q = (d * d).sum(dim=1, keepdim=True) / math.sqrt(n_classes)
    elif cons_loss_fn == 'logits_smoothl1':                                                         # :375-378
        q = F.smooth_l1_loss(logits_cons_stu, logits_cons_tea_in_stu, reduction='none').sum(dim=1, keepdim=True) / math.sqrt(n_classes)
    elif cons_loss_fn == 'bce':                                                                     # :379-382
        eps = 1e-6
        q = -(prob_cons_tea_in_stu * torch.log(prob_cons_stu + eps) +
              (1.0 - prob_cons_tea_in_stu) * torch.log(1.0 - prob_cons_stu + eps))
        q = q.sum(dim=1, keepdim=True)
    elif cons_loss_fn == 'kld':                                                                     # :383-385
        q = F.kl_div(F.log_softmax(logits_cons_stu, dim=1), prob_cons_tea_in_stu, reduction='none').sum(dim=1, keepdim=True)
    else:
        raise ValueError('Unknown consistency loss function {}'.format(cons_loss_fn))               # :386-387
    loss = (q * loss_mask).mean()                                                                   # :390
    if rampup > 0:
        loss = loss * ramp_val                                                                      # :393-394
    return loss, conf_rate


def vat_normalize_eps(x):
    """train_seg_semisup_vat_mt.py:217-220."""
    x_flat = x.view(len(x), -1)
    mag = torch.sqrt((x_flat * x_flat).sum(dim=1))
    return x / (mag[:, None, None, None] + 1e-12)


def vat_perturbation(dir_forward, x, x_hat, noise, cons_loss_fn='kld', vat_radius=0.5, adaptive_vat_radius=False):
    """`vat_direction` + `vat_perburbation`, reference train_seg_semisup_vat_mt.py:228-301 line by line.  `dir_forward(x)` is
    the direction network in eval mode (:237); `noise` is the N(0,1) draw of `normalized_noise_like` (:222), passed in so that
    the CUDA path can be driven with the same draw.  Returns x_perturb (detached), (N,3,H,W)."""
    with torch.no_grad():
        y_pred_logits = dir_forward(x).detach()                                            # :238-239
    y_pred_prob = F.softmax(y_pred_logits, dim=1)                                          # :240
    noise_scale = 1.0e-6 * x.shape[2] * x.shape[3] / 1000                                  # :243
    eps = (vat_normalize_eps(noise) * noise_scale).clone().detach().requires_grad_(True)   # :222-225
    eps_pred_logits = dir_forward(x_hat.detach() + eps)                                    # :247
    eps_pred_prob = F.softmax(eps_pred_logits, dim=1)                                      # :248
    if cons_loss_fn == 'var':                                                              # :251-253
        delta = eps_pred_prob - y_pred_prob
        loss = (delta * delta).sum()
    elif cons_loss_fn == 'bce':                                                            # :254-255
        eps_ = 1e-6
        loss = (-(y_pred_prob * torch.log(eps_pred_prob + eps_) +
                  (1.0 - y_pred_prob) * torch.log(1.0 - eps_pred_prob + eps_))).sum()
    elif cons_loss_fn == 'kld':                                                            # :256-257
        loss = F.kl_div(F.log_softmax(eps_pred_logits, dim=1), y_pred_prob, reduction='none').sum()
    elif cons_loss_fn == 'logits_var':                                                     # :258-260
        delta = eps_pred_logits - y_pred_logits
        loss = (delta * delta).sum()
    else:
        raise ValueError('Unknown consistency loss function {}'.format(cons_loss_fn))      # :261-262
    eps_adv = torch.autograd.grad(outputs=loss, inputs=eps)[0]                              # :265-268
    eps_adv_nrm = vat_normalize_eps(eps_adv)                                               # :271
    if adaptive_vat_radius:                                                                # :277-296
        delta_v = x_hat[:, :, 2:, :] - x_hat[:, :, :-2, :]
        delta_h = x_hat[:, :, :, 2:] - x_hat[:, :, :, :-2]
        delta_v = delta_v.reshape(len(delta_v), -1)
        delta_h = delta_h.reshape(len(delta_h), -1)
        adv_radius = vat_radius * torch.sqrt((delta_v ** 2).sum(dim=1) + (delta_h ** 2).sum(dim=1))[:, None, None, None] * 0.5
    else:                                                                                  # :298-299
        scale = math.sqrt(float(x_hat.shape[1] * x_hat.shape[2] * x_hat.shape[3]))
        adv_radius = vat_radius * scale
    return (eps_adv_nrm * adv_radius).detach()                                             # :301


def supervised_loss(logits, labels_n1hw):
    """nn.CrossEntropyLoss(ignore_index=255)(logits, y[:, 0]) — :126, :300."""
    return F.cross_entropy(logits, labels_n1hw[:, 0], ignore_index=255)


# ----------------------------------------------------------------------------------------------
# Bit-exact elementwise pieces (numpy float32, one rounding per operation)
# ----------------------------------------------------------------------------------------------
def ema_update(tgt, src, alpha):
    """optim_weight_ema.py:21-25 on numpy float32 arrays: t*a then + s*(1-a), three roundings."""
    a32 = np.float32(alpha)
    oma32 = np.float32(1.0 - alpha)
    return (tgt * a32).astype(np.float32) + (src * oma32).astype(np.float32)


def mix(a, b, m):
    """train_seg_semisup_mask_mt.py:350-351 (float32 numpy, separate roundings)."""
    one_minus = (np.float32(1.0) - m).astype(np.float32)
    return ((a * one_minus).astype(np.float32) + (b * m).astype(np.float32)).astype(np.float32)


def box_masks(rect_boxes, shape, invert):
    """mask_gen.py:110-116 toggle rasterisation from resolved integer boxes [y0,y1,x0,x1)."""
    n = rect_boxes.shape[0]
    out = np.zeros((n, 1) + tuple(shape), dtype=np.float32) if invert else np.ones((n, 1) + tuple(shape), dtype=np.float32)
    for i in range(n):
        for y0, y1, x0, x1 in rect_boxes[i]:
            out[i, 0, y0:y1, x0:x1] = 1 - out[i, 0, y0:y1, x0:x1]
    return out


# ----------------------------------------------------------------------------------------------
# Deterministic synthetic weights: identical in every implementation
# ----------------------------------------------------------------------------------------------
def synth_state_dict(template, seed=0, logit_gain=1.0, final_keys=()):
    """Fill a state_dict-shaped mapping {key: tensor} with well-conditioned synthetic values:
    conv weights ~ N(0, 2/fan_in); BN gamma ~ U(0.5,1.5) (U(0.15,0.35) for the last BN of a residual unit and
    for down-sample BNs so that activations stay O(1) through 33 residual units), beta ~ N(0,0.1),
    running_mean ~ N(0,0.1), running_var ~ U(0.5,1.5), biases ~ N(0,0.1); `final_keys` weights are multiplied
    by `logit_gain` so that teacher confidence straddles the threshold.  Every tensor has its own generator
    (seed, key index) so values do not depend on iteration details."""
    out = OrderedDict()
    for i, (k, v) in enumerate(template.items()):
        g = torch.Generator().manual_seed(seed * 100003 + i)
        shape = tuple(v.shape)
        if v.dtype != torch.float32:
            out[k] = torch.zeros(shape, dtype=v.dtype)
        elif k.endswith('running_mean'):
            out[k] = torch.randn(shape, generator=g) * 0.1
        elif k.endswith('running_var'):
            out[k] = torch.rand(shape, generator=g) + 0.5
        elif len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            w = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
            if k in final_keys:
                w = w * logit_gain
            out[k] = w
        elif k.endswith('.weight'):
            if k.endswith('bn3.weight') or k.endswith('downsample.1.weight'):
                out[k] = torch.rand(shape, generator=g) * 0.2 + 0.15
            else:
                out[k] = torch.rand(shape, generator=g) + 0.5
        else:
            out[k] = torch.randn(shape, generator=g) * 0.1
    return out

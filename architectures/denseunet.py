"""Drop-in for the reference's architectures/denseunet.py: U-Net on torchvision's DenseNet-161 `features` (the architecture of
BASELINE config 4: ISIC 2017, 2 classes, augmentation-driven consistency) with the H-DenseUNet style decoder shared with
architectures/resunet.py (denseunet.py:10-143).  Same `state_dict()` as the reference (1001 entries: the whole torchvision
DenseNet under `base_model.` incl. the unused ImageNet `classifier`, `line0_conv`, `decoder_blocks.0..3` in the reference's
REVERSED registration order, `final_dec_*`, `final_clf`), same attributes (BLOCK_SIZE (32, 32), MEAN / STD per instance) and
parameter-group methods, and the graph on the sm_100a kernels:

  * a dense block owns ONE (N, h, w, C0 + L*48) concatenation buffer; every layer's 3x3 convolution writes its 48 channels
    straight into its slice (no torch.cat), every layer reads a channel PREFIX of the buffer;
  * DenseNet is pre-activation: norm1 + ReLU act on that prefix (a stand-alone eval-mode BatchNorm kernel pass), conv1 (1x1)
    + norm2 + ReLU run as one fused tcgen05 GEMM launch, conv2 (3x3) is a raw GEMM into the slice;
  * transitions: stand-alone norm + ReLU, 1x1 convolution, 2x2 average pool written into the next block's buffer;
  * the gradient of a concatenation buffer is accumulated prefix by prefix (engine.Tape.contribute_slice).

Without `freeze_batchnorm()` the encoder's norms run in train mode through the batch-statistics kernels (one statistics pass
per norm over its concatenation prefix).  Input sizes must be multiples of 32 (BLOCK_SIZE), like the reference."""
from collections import OrderedDict

import numpy as np
import torch.nn as nn

from architectures.layers import B2Conv2d, B2BatchNorm2d, B2Dropout, B2Marker
from architectures.util import freeze_bn_module
from architectures import deeplab2
from architectures.resunet import DecoderBlock
from cutmix_semisup_seg_b200 import engine as E
from cutmix_semisup_seg_b200.acts import Act
from cutmix_semisup_seg_b200.netbase import B2SegNet

_DENSENET161_URL = 'https://download.pytorch.org/models/densenet161-8d451a50.pth'


class _DenseLayer(nn.Module):
    def __init__(self, num_input_features, growth_rate, bn_size):
        super(_DenseLayer, self).__init__()
        self.norm1 = B2BatchNorm2d(num_input_features)
        self.relu1 = B2Marker('relu')
        self.conv1 = B2Conv2d(num_input_features, bn_size * growth_rate, 1)
        self.norm2 = B2BatchNorm2d(bn_size * growth_rate)
        self.relu2 = B2Marker('relu')
        self.conv2 = B2Conv2d(bn_size * growth_rate, growth_rate, 3, padding=1)

    def graph(self, tape, cat, c_in):
        t = E.bn_eval_act(tape, cat.slice(0, c_in), self.norm1, relu=True)          # norm1, relu1 on the concatenated inputs
        t = E.conv_bn_act(tape, t, self.conv1, self.norm2, relu=True)               # conv1, norm2, relu2
        E.conv_bn_act(tape, t, self.conv2, out=cat.slice(c_in, self.conv2.out_channels))     # conv2 -> its slice


class _DenseBlock(nn.ModuleDict):
    def __init__(self, num_layers, num_input_features, bn_size, growth_rate):
        super(_DenseBlock, self).__init__()
        self.c_in, self.growth = num_input_features, growth_rate
        for i in range(num_layers):
            self['denselayer%d' % (i + 1)] = _DenseLayer(num_input_features + i * growth_rate, growth_rate, bn_size)
        self.c_out = num_input_features + num_layers * growth_rate

    def alloc(self, n, h, w, device):
        return Act.alloc(n, h, w, self.c_out, device, name='dense_cat')

    def graph(self, tape, cat):
        """`cat`: the block's buffer whose first c_in channels already hold the block input; returns it completed."""
        c = self.c_in
        for layer in self.values():
            layer.graph(tape, cat, c)
            c += self.growth
        return cat


class _Transition(nn.Sequential):
    def __init__(self, num_input_features, num_output_features):
        super(_Transition, self).__init__(OrderedDict([
            ('norm', B2BatchNorm2d(num_input_features)), ('relu', B2Marker('relu')),
            ('conv', B2Conv2d(num_input_features, num_output_features, 1)), ('pool', B2Marker('avgpool 2x2 s2'))]))

    def graph(self, tape, cat, out):
        t = E.bn_eval_act(tape, cat, self.norm, relu=True)
        t = E.conv_bn_act(tape, t, self.conv)
        return E.avgpool2x2(tape, t, out=out)


class TVDenseNet(nn.Module):
    """torchvision.models.densenet161 as a parameter holder (`features` + the unused `classifier`)."""

    def __init__(self, growth_rate=48, block_config=(6, 12, 36, 24), num_init_features=96, bn_size=4):
        super(TVDenseNet, self).__init__()
        feats = OrderedDict([('conv0', B2Conv2d(3, num_init_features, 7, stride=2, padding=3)),
                             ('norm0', B2BatchNorm2d(num_init_features)), ('relu0', B2Marker('relu')),
                             ('pool0', B2Marker('maxpool 3x3 s2 p1'))])
        c = num_init_features
        for i, num_layers in enumerate(block_config):
            block = _DenseBlock(num_layers, c, bn_size, growth_rate)
            feats['denseblock%d' % (i + 1)] = block
            c = block.c_out
            if i != len(block_config) - 1:
                feats['transition%d' % (i + 1)] = _Transition(c, c // 2)
                c = c // 2
        feats['norm5'] = B2BatchNorm2d(c)
        self.features = nn.Sequential(feats)
        self.classifier = nn.Linear(c, 1000)
        for m in self.modules():                       # torchvision densenet.py: kaiming_normal_ convs, unit BN, zero fc bias
            if isinstance(m, B2Conv2d):
                nn.init.kaiming_normal_(m.weight)
            elif isinstance(m, nn.Linear):
                nn.init.constant_(m.bias, 0)

    def graph(self, tape, x):
        """Returns the four taps of DenseUNet.forward (denseunet.py:100-106) and the rectified norm5 output (:108)."""
        f = self.features
        t0 = E.stem_conv(tape, x, f.conv0, f.norm0)                    # conv0, norm0, relu0 (tap 'pool0': taken before the pool)
        t = E.maxpool3x3s2(tape, t0, ceil_mode=False)                  # pool0
        taps = [t0]
        cat = f.denseblock1.alloc(t.n, t.h, t.w, t.device)
        E.copy_into(tape, t, cat.slice(0, f.denseblock1.c_in))
        for i in (1, 2, 3, 4):
            block = getattr(f, 'denseblock%d' % i)
            cat = block.graph(tape, cat)
            if i == 4:
                break
            taps.append(cat)                                            # taps 'transition1..3': the block outputs
            nxt = getattr(f, 'denseblock%d' % (i + 1))
            ncat = nxt.alloc(cat.n, cat.h // 2, cat.w // 2, cat.device)
            getattr(f, 'transition%d' % i).graph(tape, cat, ncat.slice(0, nxt.c_in))
            cat = ncat
        return taps, E.bn_eval_act(tape, cat, f.norm5, relu=True)      # norm5, then F.relu (:108)


class DenseUNet(B2SegNet):
    BLOCK_SIZE = (32, 32)
    MEAN = np.array([0.485, 0.456, 0.406])
    STD = np.array([0.229, 0.224, 0.225])

    def __init__(self, base_model, num_classes, mean, std, pretrained):
        super(DenseUNet, self).__init__()
        self.MEAN = mean
        self.STD = std
        self.pretrained = pretrained
        self.tap_names = ['pool0', 'transition1', 'transition2', 'transition3']
        self.base_model = base_model
        f = base_model.features
        enc_chn = [f.norm0.num_features, f.transition1.norm.num_features, f.transition2.norm.num_features,
                   f.transition3.norm.num_features]
        n_chn = f.norm5.num_features
        self.line0_conv = B2Conv2d(enc_chn[-1], n_chn, 1, bias=True)
        enc_chn[-1] = n_chn
        enc_chn = enc_chn[::-1]
        blocks = []
        for e_chn_a, e_chn_b in zip(enc_chn, enc_chn[1:] + enc_chn[-1:]):
            blocks.append(DecoderBlock(n_chn, e_chn_a, e_chn_b))
            n_chn = e_chn_b
        self.decoder_blocks = nn.ModuleList(blocks[::-1])              # the reference stores them in reversed order (:95)
        self.final_dec_up = B2Marker('upsample nearest x2')
        self.final_dec_conv = B2Conv2d(n_chn, 64, 3, padding=1)
        self.final_dec_drop = B2Dropout(0.3)
        self.final_dec_bn = B2BatchNorm2d(64)
        self.final_clf = B2Conv2d(64, num_classes, 1, bias=True)

    # ---- graph ---------------------------------------------------------------------------------------------------------
    def _graph_trunk(self, tape, x, in_h, in_w):
        bs = self.BLOCK_SIZE
        if in_h % bs[0] or in_w % bs[1]:
            raise ValueError('DenseUNet needs input sizes that are multiples of {} (got {}x{})'.format(bs, in_h, in_w))
        taps, out = self.base_model.graph(tape, x)
        return [(t, False) for t in taps] + [(out, True)]

    def _graph_head(self, tape, feats, in_h, in_w):
        enc_x, x = list(feats[:4]), feats[4]
        enc_x[-1] = E.conv_bn_act(tape, enc_x[-1], self.line0_conv)    # :111-112
        for dec_block, ex in zip(list(self.decoder_blocks)[::-1], enc_x[::-1]):      # :115-116
            x = dec_block.graph(tape, x, ex)
        x = E.upsample2x_add(tape, x, None)                            # :119 final_dec_up
        drop = self.final_dec_drop
        if drop.training and drop.p > 0:
            if not self.final_dec_bn.training:
                raise NotImplementedError('active dropout in front of an eval-mode BatchNorm')
            raw = E.conv_bn_act(tape, x, self.final_dec_conv)
            x = E.bn_train(tape, E.dropout_raw(tape, raw, drop), self.final_dec_bn, relu=True)
        else:
            x = E.conv_bn_act(tape, x, self.final_dec_conv, self.final_dec_bn, relu=True)       # :119-120
        c = self.final_clf
        return E.conv_bn_act(tape, x, c, ld_out=(c.out_channels + 3) // 4 * 4), False           # :121

    def _trunk_module(self):
        return self.base_model

    def forward(self, x):
        return super(DenseUNet, self).forward(x)

    # ---- reference API -------------------------------------------------------------------------------------------------
    def pretrained_parameters(self):
        if self.pretrained:
            return list(self.base_model.features.parameters())
        return []

    def new_parameters(self):
        if self.pretrained:
            pretrained_ids = set(id(p) for p in self.base_model.features.parameters())
            return [p for p in self.parameters() if id(p) not in pretrained_ids]
        return list(self.parameters())

    def freeze_batchnorm(self):
        self.base_model.apply(freeze_bn_module)


def densenet161unet(num_classes):
    return DenseUNet(TVDenseNet(), num_classes, mean=None, std=None, pretrained=False)


def densenet161unet_imagenet(num_classes):
    mean = np.array([0.485, 0.456, 0.406])
    std = np.array([0.229, 0.224, 0.225])
    base_model = TVDenseNet()
    deeplab2._load_state_into_model(base_model, deeplab2.load_pretrained_state(_DENSENET161_URL))
    return DenseUNet(base_model, num_classes, mean=mean, std=std, pretrained=True)

"""Drop-in for the reference's architectures/resunet.py: U-Net on a torchvision ResNet-50 / ResNet-101 encoder (output stride
32) with the H-DenseUNet style decoder (resunet.py:10-117).  Same `state_dict()` as the reference (the whole torchvision
ResNet under `base_model.` -- including the unused ImageNet `fc` layer --, `line0_conv`, `decoder3..0`, `final_dec_*`,
`final_clf`), same class attributes (BLOCK_SIZE (32, 32), MEAN, STD), `pretrained_parameters` / `new_parameters` /
`freeze_batchnorm` (encoder only), and the graph on the sm_100a kernels:

  * encoder: the bottleneck stack of the DeepLab backbones (tcgen05 GEMMs, frozen BN folded into the epilogue); the first
    skip connection is written as "BatchNorm output before the ReLU" in the reference (resunet.py:68-69), but the in-place
    ReLU of torchvision's ResNet rectifies that very tensor, so the fused conv-BN-ReLU stem output is the tap;
  * decoder block: nearest x2 up-sampling + skip addition in one kernel, then conv 3x3 -> train-mode BN -> ReLU;
  * final layers: up x2 -> conv 3x3 -> Dropout(0.3) on the raw conv output -> BN -> ReLU -> 1x1 classifier with bias.

Input sizes must be multiples of 32 (BLOCK_SIZE), like the reference (its skip additions fail otherwise)."""
import numpy as np
import torch.nn as nn

from architectures.layers import B2Conv2d, B2BatchNorm2d, B2Dropout, B2Marker
from architectures.util import freeze_bn_module
from architectures import deeplab2
from architectures.deeplab3plus import ResNetBackbone
from cutmix_semisup_seg_b200 import engine as E
from cutmix_semisup_seg_b200.netbase import B2SegNet

_RESNET_URLS = {'resnet50': 'https://download.pytorch.org/models/resnet50-19c8e357.pth',
                'resnet101': 'https://download.pytorch.org/models/resnet101-5d3b4d8f.pth'}


class TVResNet(ResNetBackbone):
    """torchvision.models.resnet50 / resnet101 as a parameter holder: the children the U-Net uses plus `avgpool` / `fc`,
    which the reference keeps in its state_dict although its forward never calls them."""

    def __init__(self, layers):
        super(TVResNet, self).__init__(layers, [False, False, False])
        self.avgpool = B2Marker('adaptive_avg_pool 1')
        self.fc = nn.Linear(2048, 1000)

    def graph(self, tape, x):
        # resunet.py:67-70: `r2 = x = bn1(x)` is taken before `relu`, but torchvision's ResNet uses nn.ReLU(inplace=True), so the
        # tapped tensor IS the rectified one (checked against the real module, tests/golden/net_resunet50.npz)
        r2 = E.stem_conv(tape, x, self.conv1, self.bn1)
        t = E.maxpool3x3s2(tape, r2, ceil_mode=False)
        feats = [r2]
        for name in ('layer1', 'layer2', 'layer3', 'layer4'):              # :73-76
            for unit in getattr(self, name):
                t = unit.graph(tape, t)
            feats.append(t)
        return feats


class DecoderBlock(nn.Module):
    """Reference resunet.py:10-34."""

    def __init__(self, x_chn_in, skip_chn_in, chn_out):
        super(DecoderBlock, self).__init__()
        if x_chn_in != skip_chn_in:
            raise ValueError('x_chn_in != skip_chn_in')
        self.x_chn_in, self.skip_chn_in, self.chn_out = x_chn_in, skip_chn_in, chn_out
        self.up = B2Marker('upsample nearest x2')
        self.conv = B2Conv2d(x_chn_in, chn_out, 3, padding=1)
        self.conv_bn = B2BatchNorm2d(chn_out)

    def graph(self, tape, x_in, skip_in):
        if x_in.c != self.x_chn_in:
            raise ValueError('x_in.shape[1]={}, self.x_chn_in={}'.format(x_in.c, self.x_chn_in))
        if skip_in.c != self.skip_chn_in:
            raise ValueError('skip_in.shape[1]={}, self.skip_chn_in={}'.format(skip_in.c, self.skip_chn_in))
        x = E.upsample2x_add(tape, x_in, skip_in)                          # :31-32
        return E.conv_bn_act(tape, x, self.conv, self.conv_bn, relu=True)  # :33


class ResUNet(B2SegNet):
    BLOCK_SIZE = (32, 32)
    MEAN = np.array([0.485, 0.456, 0.406])
    STD = np.array([0.229, 0.224, 0.225])

    def __init__(self, base_model, num_classes, pretrained):
        super(ResUNet, self).__init__()
        self.base_model = base_model
        self.pretrained = pretrained
        self.line0_conv = B2Conv2d(2048, 1024, 1, bias=True)
        self.decoder3 = DecoderBlock(1024, 1024, 512)
        self.decoder2 = DecoderBlock(512, 512, 256)
        self.decoder1 = DecoderBlock(256, 256, 64)
        self.decoder0 = DecoderBlock(64, 64, 64)
        self.final_dec_up = B2Marker('upsample nearest x2')
        self.final_dec_conv = B2Conv2d(64, 64, 3, padding=1)
        self.final_dec_drop = B2Dropout(0.3)
        self.final_dec_bn = B2BatchNorm2d(64)
        self.final_clf = B2Conv2d(64, num_classes, 1, bias=True)

    # ---- graph ---------------------------------------------------------------------------------------------------------
    def _graph_trunk(self, tape, x, in_h, in_w):
        bs = self.BLOCK_SIZE
        if in_h % bs[0] or in_w % bs[1]:
            raise ValueError('ResUNet needs input sizes that are multiples of {} (got {}x{})'.format(bs, in_h, in_w))
        r2, r4, r8, r16, r32 = self.base_model.graph(tape, x)
        # every tap but the last also feeds the next encoder stage
        return [(r2, False), (r4, False), (r8, False), (r16, False), (r32, True)]

    def _graph_head(self, tape, feats, in_h, in_w):
        r2, r4, r8, r16, r32 = feats
        x = E.conv_bn_act(tape, r32, self.line0_conv)                      # :79
        x = self.decoder3.graph(tape, x, r16)                              # :82-85
        x = self.decoder2.graph(tape, x, r8)
        x = self.decoder1.graph(tape, x, r4)
        x = self.decoder0.graph(tape, x, r2)
        x = E.upsample2x_add(tape, x, None)                                # :88 final_dec_up
        drop = self.final_dec_drop
        if drop.training and drop.p > 0:
            if not self.final_dec_bn.training:
                raise NotImplementedError('active dropout in front of an eval-mode BatchNorm')
            raw = E.conv_bn_act(tape, x, self.final_dec_conv)              # conv, then Dropout on the raw output, then BN + ReLU
            x = E.bn_train(tape, E.dropout_raw(tape, raw, drop), self.final_dec_bn, relu=True)
        else:
            x = E.conv_bn_act(tape, x, self.final_dec_conv, self.final_dec_bn, relu=True)       # :88-89
        c = self.final_clf
        logits = E.conv_bn_act(tape, x, c, ld_out=(c.out_channels + 3) // 4 * 4)               # :90
        return logits, False          # already at the input resolution: the final resize is the identity

    def _trunk_module(self):
        return self.base_model

    def forward(self, x):
        return super(ResUNet, self).forward(x)

    # ---- reference API -------------------------------------------------------------------------------------------------
    def pretrained_parameters(self):
        if self.pretrained:
            return list(self.base_model.parameters())
        return []

    def new_parameters(self):
        if self.pretrained:
            pretrained_ids = set(id(p) for p in self.base_model.parameters())
            return [p for p in self.parameters() if id(p) not in pretrained_ids]
        return list(self.parameters())

    def freeze_batchnorm(self):
        self.base_model.apply(freeze_bn_module)


def _tv_resnet(name, layers, pretrained):
    net = TVResNet(layers)
    if pretrained:
        deeplab2._load_state_into_model(net, deeplab2.load_pretrained_state(_RESNET_URLS[name]))
    return net


def resnet50unet(num_classes, pretrained=True):
    return ResUNet(_tv_resnet('resnet50', [3, 4, 6, 3], pretrained), num_classes, pretrained=pretrained)


def resnet101unet(num_classes, pretrained=True):
    return ResUNet(_tv_resnet('resnet101', [3, 4, 23, 3], pretrained), num_classes, pretrained=pretrained)

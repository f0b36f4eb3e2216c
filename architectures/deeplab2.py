"""Drop-in for the reference's architectures/deeplab2.py: DeepLab v2 with a ResNet-101 backbone.

Same factory functions, constructor arguments, attributes (`BLOCK_SIZE`, `MEAN`, `STD`), methods
(`pretrained_parameters`, `new_parameters`, `freeze_batchnorm`) and `state_dict()` keys / shapes /
dtypes (632 keys) as the reference (deeplab2.py:131-307), but the forward/backward arithmetic runs in
the sm_100a kernels of libb200seg.so through cutmix_semisup_seg_b200.engine:

  * every conv + frozen BatchNorm + residual + ReLU is ONE tcgen05 implicit-GEMM launch,
  * layer5 (two dilated 2048->C convolutions, deeplab2.py:124-128) is two GEMM launches of which the
    second accumulates onto the first through its epilogue,
  * the final `align_corners=True` bilinear resize (deeplab2.py:204) writes NCHW logits directly.

Reference quirks that are reproduced on purpose (SURVEY.md §7 "hard parts"):
  * `Classifier_Module.forward` returns inside its loop, so only dilations 6 and 12 are used although
    four convolutions exist in the state dict (deeplab2.py:124-128);
  * `pretrained_parameters()` walks `modules()` x `parameters()` and therefore yields most
    backbone weights several times (deeplab2.py:208-230): the optimiser sees 314 entries / 104
    unique tensors;
  * BatchNorm affine parameters have `requires_grad=False` (deeplab2.py:72-84), stride sits on the
    first 1x1 convolution of a bottleneck (deeplab2.py:70), the max-pool uses `ceil_mode=True`
    (deeplab2.py:146).
"""
import os

import numpy as np
import torch
import torch.nn as nn

from architectures.layers import B2Conv2d, B2BatchNorm2d, B2Marker
from architectures.util import freeze_bn_module
from cutmix_semisup_seg_b200 import engine as E
from cutmix_semisup_seg_b200.acts import Act
from cutmix_semisup_seg_b200.netbase import B2SegNet

_RESNET_101_DEEPLAB_COCO_URL = 'http://vllab1.ucmerced.edu/~whung/adv-semi-seg/resnet101COCO-41f33a49.pth'
_RESNET_101_IMAGENET_URL = 'https://download.pytorch.org/models/resnet101-5d3b4d8f.pth'


def _frozen_bn(channels):
    bn = B2BatchNorm2d(channels)
    for p in bn.parameters():
        p.requires_grad = False
    return bn


class Bottleneck(nn.Module):
    """1x1 (stride) -> 3x3 (dilated) -> 1x1 (x4) residual unit, reference deeplab2.py:65-109."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None):
        super(Bottleneck, self).__init__()
        self.conv1 = B2Conv2d(inplanes, planes, 1, stride=stride)
        self.bn1 = _frozen_bn(planes)
        self.conv2 = B2Conv2d(planes, planes, 3, stride=1, padding=dilation, dilation=dilation)
        self.bn2 = _frozen_bn(planes)
        self.conv3 = B2Conv2d(planes, planes * 4, 1)
        self.bn3 = _frozen_bn(planes * 4)
        self.relu = B2Marker('relu')
        self.downsample = downsample
        self.stride = stride

    def graph(self, tape, x):
        t = E.conv_bn_act(tape, x, self.conv1, self.bn1, relu=True)
        t = E.conv_bn_act(tape, t, self.conv2, self.bn2, relu=True)
        if self.downsample is not None:
            res = E.conv_bn_act(tape, x, self.downsample[0], self.downsample[1], relu=False)
        else:
            res = x
        return E.conv_bn_act(tape, t, self.conv3, self.bn3, residual=res, relu=True)


class Classifier_Module(nn.Module):
    """ASPP-style classifier of DeepLab v2 (reference deeplab2.py:112-128)."""

    def __init__(self, dilation_series, padding_series, num_classes):
        super(Classifier_Module, self).__init__()
        self.conv2d_list = nn.ModuleList()
        for dilation, padding in zip(dilation_series, padding_series):
            self.conv2d_list.append(B2Conv2d(2048, num_classes, 3, stride=1, padding=padding, dilation=dilation,
                                             bias=True))
        for m in self.conv2d_list:
            m.weight.data.normal_(0, 0.01)

    def graph(self, tape, x):
        # Only the first two branches contribute: the reference returns from inside its accumulation loop.
        ld = (self.conv2d_list[0].out_channels + 3) // 4 * 4
        first = E.conv_bn_act(tape, x, self.conv2d_list[0], ld_out=ld)
        if len(self.conv2d_list) < 2:
            return first
        return E.conv_bn_act(tape, x, self.conv2d_list[1], residual=first, ld_out=ld)


class ResNetDeepLab(B2SegNet):
    BLOCK_SIZE = (1, 1)

    def __init__(self, block, layers, num_classes, mean, std):
        super(ResNetDeepLab, self).__init__()
        self.MEAN = mean
        self.STD = std
        self.inplanes = 64
        self.conv1 = B2Conv2d(3, 64, 7, stride=2, padding=3)
        self.bn1 = _frozen_bn(64)
        self.relu = B2Marker('relu')
        self.maxpool = B2Marker('maxpool 3x3 s2 p1 ceil_mode=True')
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=1, dilation=2)
        self.layer4 = self._make_layer(block, 512, layers[3], stride=1, dilation=4)
        self.layer5 = Classifier_Module([6, 12, 18, 24], [6, 12, 18, 24], num_classes)
        for m in self.modules():
            if isinstance(m, B2Conv2d):
                m.weight.data.normal_(0, 0.01)        # reference deeplab2.py:153-156
            elif isinstance(m, B2BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _make_layer(self, block, planes, blocks, stride=1, dilation=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion or dilation in (2, 4):
            downsample = nn.Sequential(B2Conv2d(self.inplanes, planes * block.expansion, 1, stride=stride),
                                       _frozen_bn(planes * block.expansion))
        units = [block(self.inplanes, planes, stride, dilation=dilation, downsample=downsample)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            units.append(block(self.inplanes, planes, dilation=dilation))
        return nn.Sequential(*units)

    # ---- engine graph (reference forward: deeplab2.py:183-206) -------------------------------
    def _graph_trunk(self, tape, x, in_h, in_w):
        # every BatchNorm of DeepLab v2 is a frozen one: the whole network is the batch-invariant trunk and the "head"
        # is the final resize of the low-resolution logits
        t = E.stem_conv(tape, x, self.conv1, self.bn1)
        t = E.maxpool3x3s2(tape, t, ceil_mode=True)
        for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
            for unit in layer:
                t = unit.graph(tape, t)
        t = self.layer5.graph(tape, t)
        return [(t, True)]

    def _graph_head(self, tape, feats, in_h, in_w):
        return feats[0], True

    def _trunk_module(self):
        return self

    def forward(self, x, use_dropout=False):
        return super(ResNetDeepLab, self).forward(x)

    # ---- optimiser parameter groups ------------------------------------------------------------
    def pretrained_parameters(self):
        """Backbone parameters for the 0.1 x lr group.  The traversal order and the repetitions of the
        reference generator (one visit per enclosing module) are reproduced: torch.optim sees the same
        list, so a tensor listed k times is stepped k times per `optimizer.step()`."""
        for top in (self.conv1, self.bn1, self.layer1, self.layer2, self.layer3, self.layer4):
            for sub in top.modules():
                for p in sub.parameters():
                    if p.requires_grad:
                        yield p

    def new_parameters(self):
        for p in self.layer5.parameters():
            yield p

    def freeze_batchnorm(self):
        self.apply(freeze_bn_module)


# --------------------------------------------------------------------------------------------- weights
def _find_cached(url):
    name = os.path.basename(url)
    roots = [os.environ.get('B200SEG_WEIGHTS', ''), os.path.join(torch.hub.get_dir(), 'checkpoints')]
    for r in roots:
        if r and os.path.exists(os.path.join(r, name)):
            return os.path.join(r, name)
    return None


def load_pretrained_state(url):
    """Pretrained weights from a local cache (this build runs offline): `$B200SEG_WEIGHTS/<file>` or the
    torch hub checkpoint directory; downloads through torch.hub when a network is available."""
    path = _find_cached(url)
    if path is not None:
        return torch.load(path, map_location='cpu')
    try:
        return torch.hub.load_state_dict_from_url(url, map_location='cpu')
    except Exception as e:  # offline
        raise RuntimeError('pretrained weights {} are not cached locally and cannot be downloaded ({}); pass '
                           'pretrained=False or put the file under $B200SEG_WEIGHTS'.format(url, e))


def _load_state_into_model(model, state_dict, verbose=False):
    """Copy every entry whose name and shape match (reference deeplab2.py:310-322)."""
    own = model.state_dict()
    for name, dst in own.items():
        src = state_dict.get(name)
        if src is None:
            if verbose:
                print('Could not find {}'.format(name))
        elif tuple(src.shape) == tuple(dst.shape):
            dst.copy_(src)
        elif verbose:
            print('{} -> {}'.format(tuple(src.shape), tuple(dst.shape)))
    return model


_IMAGENET_MEAN = np.array([0.485, 0.456, 0.406])
_IMAGENET_STD = np.array([0.229, 0.224, 0.225])
# Hung et al. normalisation: BGR ImageNet mean on the 0..255 scale, no std scaling (deeplab2.py:249-266)
_HUNG_MEAN = np.array((104.00698793, 116.66876762, 122.67891434))[::-1] / 255.0
_HUNG_STD = np.array([1, 1, 1]) / 255.0


def resnet101_deeplab_coco(num_classes=21, pretrained=True):
    model = ResNetDeepLab(Bottleneck, [3, 4, 23, 3], num_classes, _HUNG_MEAN, _HUNG_STD)
    if pretrained:
        _load_state_into_model(model, load_pretrained_state(_RESNET_101_DEEPLAB_COCO_URL))
    return model


def resnet101_deeplab_imagenet(num_classes=21, pretrained=True):
    model = ResNetDeepLab(Bottleneck, [3, 4, 23, 3], num_classes, _IMAGENET_MEAN, _IMAGENET_STD)
    if pretrained:
        _load_state_into_model(model, load_pretrained_state(_RESNET_101_IMAGENET_URL))
    return model


def resnet101_deeplab_imagenet_mittal_std(num_classes=21, pretrained=True):
    model = ResNetDeepLab(Bottleneck, [3, 4, 23, 3], num_classes, _HUNG_MEAN, _HUNG_STD)
    if pretrained:
        _load_state_into_model(model, load_pretrained_state(_RESNET_101_IMAGENET_URL))
    return model

"""Drop-in for the reference's architectures/deeplab3plus.py: DeepLab v3+ (ResNet-101, output stride 8), and the DeepLab v3
model of torchvision that the reference wraps with the same `DeepLabv3Wrapper` (network_architectures.py:75-98).

The reference composes torchvision's ResNet-101 (`replace_stride_with_dilation=[False, True, True]`),
`IntermediateLayerGetter`, torchvision's `ASPP` and its own `DeepLabHeadV3Plus`
(deeplab3plus.py:26-101).  This module reproduces the same `state_dict()` (680 keys: 227 conv weights,
bias of the last conv, 113 x {weight, bias, running_mean, running_var, num_batches_tracked}), the
wrapper API (`DeepLabv3Wrapper`: BLOCK_SIZE, MEAN, STD, forward(x, feature_maps=False,
use_dropout=False), freeze_batchnorm (backbone only), pretrained_parameters / new_parameters) and runs
the graph on the sm_100a kernels:

  * backbone bottlenecks: conv + (frozen) BN + ReLU (+ residual) per tcgen05 GEMM launch;
  * ASPP: the five branches write their BN/ReLU outputs straight into channel slices of one
    (N,64,64,1280) buffer (no torch.cat); dilation-12/24/36 taps that fall entirely into the zero
    padding are skipped by the GEMM's tile scheduler;
  * decoder: project (48 ch) and the x2 bilinear up-sampled ASPP output (256 ch) share one
    (N,128,128,304) buffer; train-mode BatchNorm of the head = statistics kernel + fused apply;
  * final x4 `align_corners=False` resize writes NCHW logits.
"""
import numpy as np
import torch
import torch.nn as nn

from architectures.layers import B2Conv2d, B2BatchNorm2d, B2Dropout, B2Marker
from architectures.util import freeze_bn_module
from architectures import deeplab2
from cutmix_semisup_seg_b200 import engine as E
from cutmix_semisup_seg_b200.acts import Act
from cutmix_semisup_seg_b200.netbase import B2SegNet


class TVBottleneck(nn.Module):
    """torchvision ResNet bottleneck (stride and dilation on the 3x3 convolution)."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, dilation=1):
        super(TVBottleneck, self).__init__()
        self.conv1 = B2Conv2d(inplanes, planes, 1)
        self.bn1 = B2BatchNorm2d(planes)
        self.conv2 = B2Conv2d(planes, planes, 3, stride=stride, padding=dilation, dilation=dilation)
        self.bn2 = B2BatchNorm2d(planes)
        self.conv3 = B2Conv2d(planes, planes * 4, 1)
        self.bn3 = B2BatchNorm2d(planes * 4)
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


class ResNetBackbone(nn.Module):
    """The children of torchvision's resnet101 that IntermediateLayerGetter keeps (conv1 .. layer4)."""

    def __init__(self, layers, replace_stride_with_dilation):
        super(ResNetBackbone, self).__init__()
        self.inplanes = 64
        self.dilation = 1
        self.conv1 = B2Conv2d(3, 64, 7, stride=2, padding=3)
        self.bn1 = B2BatchNorm2d(64)
        self.relu = B2Marker('relu')
        self.maxpool = B2Marker('maxpool 3x3 s2 p1')
        self.layer1 = self._make_layer(64, layers[0])
        self.layer2 = self._make_layer(128, layers[1], stride=2, dilate=replace_stride_with_dilation[0])
        self.layer3 = self._make_layer(256, layers[2], stride=2, dilate=replace_stride_with_dilation[1])
        self.layer4 = self._make_layer(512, layers[3], stride=2, dilate=replace_stride_with_dilation[2])
        del self.inplanes, self.dilation
        for m in self.modules():
            if isinstance(m, B2Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')

    def _make_layer(self, planes, blocks, stride=1, dilate=False):
        previous_dilation = self.dilation
        if dilate:
            self.dilation *= stride
            stride = 1
        downsample = None
        if stride != 1 or self.inplanes != planes * 4:
            downsample = nn.Sequential(B2Conv2d(self.inplanes, planes * 4, 1, stride=stride), B2BatchNorm2d(planes * 4))
        units = [TVBottleneck(self.inplanes, planes, stride, downsample, previous_dilation)]
        self.inplanes = planes * 4
        for _ in range(1, blocks):
            units.append(TVBottleneck(self.inplanes, planes, dilation=self.dilation))
        return nn.Sequential(*units)

    def graph(self, tape, x):
        t = E.stem_conv(tape, x, self.conv1, self.bn1)
        t = E.maxpool3x3s2(tape, t, ceil_mode=False)
        feats = {}
        for name in ('layer1', 'layer2', 'layer3', 'layer4'):
            for unit in getattr(self, name):
                t = unit.graph(tape, t)
            feats[name] = t
        return {'low_level': feats['layer1'], 'out': feats['layer4']}


def _conv_bn_relu(cin, cout, k, dilation=1):
    return nn.Sequential(B2Conv2d(cin, cout, k, padding=0 if k == 1 else dilation, dilation=dilation),
                         B2BatchNorm2d(cout), B2Marker('relu'))


class ASPP(nn.Module):
    """torchvision.models.segmentation.deeplabv3.ASPP: 1x1 + three dilated 3x3 + image pooling,
    concat (5 x 256) -> 1x1 -> BN -> ReLU -> Dropout(0.5)."""

    def __init__(self, in_channels, atrous_rates, out_channels=256):
        super(ASPP, self).__init__()
        branches = [_conv_bn_relu(in_channels, out_channels, 1)]
        for rate in atrous_rates:
            branches.append(_conv_bn_relu(in_channels, out_channels, 3, dilation=rate))
        branches.append(nn.Sequential(B2Marker('adaptive_avg_pool 1'), B2Conv2d(in_channels, out_channels, 1),
                                      B2BatchNorm2d(out_channels), B2Marker('relu')))
        self.convs = nn.ModuleList(branches)
        self.project = nn.Sequential(B2Conv2d(len(self.convs) * out_channels, out_channels, 1),
                                     B2BatchNorm2d(out_channels), B2Marker('relu'), B2Dropout(0.5))
        self.out_channels = out_channels

    def graph(self, tape, x):
        oc = self.out_channels
        nb = len(self.convs)
        cat = Act.alloc(x.n, x.h, x.w, nb * oc, x.device, name='aspp_cat')
        for i in range(nb - 1):
            seq = self.convs[i]
            E.conv_bn_act(tape, x, seq[0], seq[1], relu=True, out=cat.slice(i * oc, oc))
        pool = self.convs[nb - 1]
        v = E.global_avg_pool(tape, x)
        v = E.conv_bn_act(tape, v, pool[1], pool[2], relu=True)
        E.broadcast(tape, v, cat.slice((nb - 1) * oc, oc))       # bilinear from 1x1 == broadcast
        return E.conv_bn_act(tape, cat, self.project[0], self.project[1], relu=True, dropout=self.project[3])


class DeepLabHeadV3Plus(nn.Module):
    """Reference deeplab3plus.py:26-64."""

    def __init__(self, in_channels, low_level_channels, num_classes, aspp_dilate=(12, 24, 36)):
        super(DeepLabHeadV3Plus, self).__init__()
        self.project = _conv_bn_relu(low_level_channels, 48, 1)
        self.aspp = ASPP(in_channels, aspp_dilate)
        self.classifier = nn.Sequential(
            B2Conv2d(304, 256, 3, padding=1), B2BatchNorm2d(256), B2Marker('relu'),
            B2Conv2d(256, 256, 3, padding=1), B2BatchNorm2d(256), B2Marker('relu'),
            B2Conv2d(256, num_classes, 1, bias=True))
        for m in self.modules():
            if isinstance(m, B2Conv2d):
                nn.init.kaiming_normal_(m.weight)

    def graph(self, tape, feature):
        low = feature['low_level']
        cat = Act.alloc(low.n, low.h, low.w, 304, low.device, name='decoder_cat')
        E.conv_bn_act(tape, low, self.project[0], self.project[1], relu=True, out=cat.slice(0, 48))
        a = self.aspp.graph(tape, feature['out'])
        E.bilinear(tape, a, low.h, low.w, False, out=cat.slice(48, 256))
        c = self.classifier
        t = E.conv_bn_act(tape, cat, c[0], c[1], relu=True)
        t = E.conv_bn_act(tape, t, c[3], c[4], relu=True)
        return E.conv_bn_act(tape, t, c[6], ld_out=(c[6].out_channels + 3) // 4 * 4)


class DeepLabHead(nn.Sequential):
    """torchvision.models.segmentation.deeplabv3.DeepLabHead (the DeepLab v3 head the reference wraps for its
    `resnet101_deeplabv3_*` architectures, network_architectures.py:75-98): ASPP -> 3x3 conv -> BN -> ReLU -> 1x1 conv.
    A Sequential like torchvision's, so the state_dict keys are `classifier.0.convs...`, `classifier.1.weight`, ..."""

    def __init__(self, in_channels, num_classes, aspp_dilate=(12, 24, 36)):
        super(DeepLabHead, self).__init__(
            ASPP(in_channels, aspp_dilate),
            B2Conv2d(256, 256, 3, padding=1), B2BatchNorm2d(256), B2Marker('relu'),
            B2Conv2d(256, num_classes, 1, bias=True))

    def graph(self, tape, feature):
        a = self[0].graph(tape, feature['out'])
        t = E.conv_bn_act(tape, a, self[1], self[2], relu=True)
        return E.conv_bn_act(tape, t, self[4], ld_out=(self[4].out_channels + 3) // 4 * 4)


class DeepLabV3(nn.Module):
    """torchvision.models.segmentation.DeepLabV3 without auxiliary classifier (`deeplabv3_resnet101(pretrained=False,
    num_classes=...)`: aux_loss is off unless the COCO weights are requested)."""

    def __init__(self, backbone, classifier):
        super(DeepLabV3, self).__init__()
        self.backbone = backbone
        self.classifier = classifier
        self.aux_classifier = None


class DeepLabV3Plus(nn.Module):
    def __init__(self, backbone, classifier):
        super(DeepLabV3Plus, self).__init__()
        self.backbone = backbone
        self.classifier = classifier


def _deeplabv3plus(backbone_name, num_classes, output_stride, pretrained_backbone):
    if backbone_name != 'resnet101':
        raise NotImplementedError('only the resnet101 backbone is built for B200')
    if output_stride == 8:
        replace_stride_with_dilation, aspp_dilate = [False, True, True], [12, 24, 36]
    else:
        replace_stride_with_dilation, aspp_dilate = [False, False, True], [6, 12, 18]
    backbone = ResNetBackbone([3, 4, 23, 3], replace_stride_with_dilation)
    if pretrained_backbone:
        deeplab2._load_state_into_model(backbone, deeplab2.load_pretrained_state(deeplab2._RESNET_101_IMAGENET_URL))
    return DeepLabV3Plus(backbone, DeepLabHeadV3Plus(2048, 256, num_classes, aspp_dilate))


class DeepLabv3Wrapper(B2SegNet):
    BLOCK_SIZE = (1, 1)
    MEAN = np.array([0.485, 0.456, 0.406])
    STD = np.array([0.229, 0.224, 0.225])

    def __init__(self, model, pretraining=None):
        super(DeepLabv3Wrapper, self).__init__()
        self.deeplab = model
        self.pretraining = pretraining

    def _graph_trunk(self, tape, x, in_h, in_w):
        feats = self.deeplab.backbone.graph(tape, x)
        if isinstance(self.deeplab.classifier, DeepLabHead):     # DeepLab v3: the head reads layer4 only
            return [(feats['out'], True)]
        # layer1's output also feeds layer2; layer4's output is consumed by the head only
        return [(feats['low_level'], False), (feats['out'], True)]

    def _graph_head(self, tape, feats, in_h, in_w):
        if isinstance(self.deeplab.classifier, DeepLabHead):
            return self.deeplab.classifier.graph(tape, {'out': feats[0]}), False
        return self.deeplab.classifier.graph(tape, {'low_level': feats[0], 'out': feats[1]}), False

    def _trunk_module(self):
        return self.deeplab.backbone

    def forward(self, x, feature_maps=False, use_dropout=False):
        return super(DeepLabv3Wrapper, self).forward(x)

    def freeze_batchnorm(self):
        self.deeplab.backbone.apply(freeze_bn_module)

    def _backbone_parameters(self):
        return list(self.deeplab.backbone.parameters())

    def _classifier_end_parameters(self):
        if isinstance(self.deeplab.classifier, DeepLabHead):
            return list(self.deeplab.classifier[-1].parameters())
        if isinstance(self.deeplab.classifier, DeepLabHeadV3Plus):
            return list(self.deeplab.classifier.classifier[-1].parameters())
        raise TypeError('Oh dear, seem to have encountered unknown classifier head type {}'.format(
            type(self.deeplab.classifier)))

    def pretrained_parameters(self):
        if self.pretraining is None:
            return []
        if self.pretraining == 'imagenet':
            return self._backbone_parameters()
        if self.pretraining == 'coco':
            new_ids = set(id(p) for p in self._classifier_end_parameters())
            return [p for p in self.parameters() if id(p) not in new_ids]
        raise ValueError('Unknown pretraining {}'.format(self.pretraining))

    def new_parameters(self):
        if self.pretraining is None:
            return list(self.parameters())
        if self.pretraining == 'imagenet':
            backbone_ids = set(id(p) for p in self._backbone_parameters())
            return [p for p in self.parameters() if id(p) not in backbone_ids]
        if self.pretraining == 'coco':
            return self._classifier_end_parameters()
        raise ValueError('Unknown pretraining {}'.format(self.pretraining))


def resnet101_deeplabv3plus_imagenet(num_classes, pretrained=True):
    deeplab = _deeplabv3plus('resnet101', num_classes, 8, pretrained)
    return DeepLabv3Wrapper(deeplab)


_DEEPLABV3_COCO_URL = 'https://download.pytorch.org/models/deeplabv3_resnet101_coco-586e9e4e.pth'


def _deeplabv3(num_classes):
    """torchvision's `deeplabv3_resnet101(pretrained=False, num_classes=num_classes)`: ResNet-101 with the strides of layer3 /
    layer4 replaced by dilation (output stride 8), DeepLabHead, no auxiliary classifier."""
    backbone = ResNetBackbone([3, 4, 23, 3], [False, True, True])
    return DeepLabV3(backbone, DeepLabHead(2048, num_classes))


def resnet101_deeplabv3_coco(num_classes=21, pretrained=True):
    """Reference network_architectures.py:75-84 (the wrapper is built WITHOUT a `pretraining` tag there, so every parameter
    is a "new" parameter: kept)."""
    deeplab = _deeplabv3(num_classes)
    if pretrained:
        deeplab2._load_state_into_model(deeplab, deeplab2.load_pretrained_state(_DEEPLABV3_COCO_URL))
    return DeepLabv3Wrapper(deeplab)


def resnet101_deeplabv3_imagenet(num_classes=21, pretrained=True):
    """Reference network_architectures.py:87-98: the DeepLab v2 ImageNet ResNet-101 weights under a `backbone.` prefix."""
    deeplab = _deeplabv3(num_classes)
    if pretrained:
        state = deeplab2.load_pretrained_state(deeplab2._RESNET_101_IMAGENET_URL)
        deeplab2._load_state_into_model(deeplab, {'backbone.{}'.format(k): v for k, v in state.items()})
    return DeepLabv3Wrapper(deeplab)

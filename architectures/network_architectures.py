"""Drop-in for the reference's architectures/network_architectures.py: the `seg` architecture registry
(`seg.get(name)(num_classes, pretrained=...)`), plus `robust_binary_crossentropy` and
`sigmoid_rampup` (network_architectures.py:15-130).

The architectures on the B200 hot path are built natively (DeepLab v2 and DeepLab v3+ on ResNet-101), and so are
torchvision's DeepLab v3, which shares every layer type with them, the ResNet U-Nets (architectures/resunet.py)
and the DenseNet-161 U-Net (architectures/denseunet.py).  The other names of the reference registry stay registered so that `seg.names()` matches,
but constructing them raises NotImplementedError (they are outside BASELINE.json's north_star).
"""
import sys

import numpy as np
import torch

from architectures import deeplab2, deeplab3plus, denseunet, resunet


class ArchRegistry(object):
    """name -> constructor registry with a decorator interface."""

    def __init__(self):
        self.archs = {}

    def register(self, name):
        def deco(arch):
            self.archs[name] = arch
            return arch
        return deco

    def get(self, name):
        return self.archs[name]

    def names(self):
        return self.archs.keys()


seg = ArchRegistry()


def _not_built(name):
    def ctor(*args, **kwargs):
        raise NotImplementedError('architecture {!r} is not part of the B200 hot path (only the resnet101_deeplab* '
                                  'architectures are built natively)'.format(name))
    ctor.__name__ = name
    return ctor


for _name in ('resnet101_pspnet_imagenet',):
    seg.register(_name)(_not_built(_name))


@seg.register('resnet50unet_imagenet')
def resnet50unet_imagenet(num_classes, pretrained=True):
    return resunet.resnet50unet(num_classes, pretrained=pretrained)


@seg.register('resnet101unet_imagenet')
def resnet101unet_imagenet(num_classes, pretrained=True):
    return resunet.resnet101unet(num_classes, pretrained=pretrained)


@seg.register('densenet161unet')
def densenet161unet(num_classes):
    return denseunet.densenet161unet(num_classes)


@seg.register('densenet161unet_imagenet')
def densenet161unet_imagenet(num_classes):
    return denseunet.densenet161unet_imagenet(num_classes)


@seg.register('resnet101_deeplab_coco')
def resnet101_deeplab_coco(num_classes=21, pretrained=True):
    return deeplab2.resnet101_deeplab_coco(num_classes=num_classes, pretrained=pretrained)


@seg.register('resnet101_deeplab_imagenet')
def resnet101_deeplab_imagenet(num_classes=21, pretrained=True):
    return deeplab2.resnet101_deeplab_imagenet(num_classes=num_classes, pretrained=pretrained)


@seg.register('resnet101_deeplab_imagenet_mittal_std')
def resnet101_deeplab_imagenet_mittal_std(num_classes=21, pretrained=True):
    return deeplab2.resnet101_deeplab_imagenet_mittal_std(num_classes=num_classes, pretrained=pretrained)


@seg.register('resnet101_deeplabv3_coco')
def resnet101_deeplabv3_coco(num_classes=21, pretrained=True):
    return deeplab3plus.resnet101_deeplabv3_coco(num_classes=num_classes, pretrained=pretrained)


@seg.register('resnet101_deeplabv3_imagenet')
def resnet101_deeplabv3_imagenet(num_classes=21, pretrained=True):
    return deeplab3plus.resnet101_deeplabv3_imagenet(num_classes=num_classes, pretrained=pretrained)


@seg.register('resnet101_deeplabv3plus_imagenet')
def resnet101_deeplabv3plus_imagenet(num_classes=21, pretrained=True):
    return deeplab3plus.resnet101_deeplabv3plus_imagenet(num_classes=num_classes, pretrained=pretrained)


def robust_binary_crossentropy(pred, tgt, eps=1e-6):
    """Element-wise BCE with an epsilon inside both logs (used by the 'bce' consistency loss)."""
    return -(tgt * torch.log(pred + eps) + (1.0 - tgt) * torch.log(1.0 - pred + eps))


EPS = sys.float_info.epsilon


def sigmoid_rampup(current, rampup_length):
    """exp(-5 (1 - t/T)^2) ramp of Laine & Aila (https://arxiv.org/abs/1610.02242)."""
    if rampup_length == 0:
        return 1.0
    t = float(np.clip(current, 0.0, rampup_length))
    phase = 1.0 - t / rampup_length
    return float(np.exp(-5.0 * phase * phase))

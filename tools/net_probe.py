"""GPU parity probe of whole networks against the torch CPU oracle (development tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from collections import OrderedDict
import torch
from architectures import network_architectures as na
from oracle import torch_oracle as TO

dev = torch.device('cuda:0')


def run(kind, N, H, W, C, freeze, precisions=('3xtf32', 'tf32')):
    torch.manual_seed(0)
    net = na.seg.get(kind)(C, pretrained=False)
    sd = TO.synth_state_dict(net.state_dict(), seed=1)
    x = torch.randn(N, 3, H, W)
    dm = (torch.rand(N, -(-H // 8), -(-W // 8), 256) > 0.5).float()
    # oracle fp64
    sd64 = OrderedDict((k, v.double().clone() if v.dtype == torch.float32 else v.clone()) for k, v in sd.items())
    for k, p in net.named_parameters():
        if p.requires_grad:
            sd64[k].requires_grad_(True)
    t = time.time()
    if 'v3plus' in kind:
        yo = TO.deeplab3plus_forward(sd64, x.double(), backbone_bn_train=not freeze, head_bn_train=True,
                                     dropout_masks=[dm.permute(0, 3, 1, 2).double()])
    else:
        yo = TO.deeplab2_forward(sd64, x.double(), bn_train=not freeze)
    dy = torch.randn(yo.shape)
    yo.backward(dy.double())
    print('{} N{} {}x{} C{} freeze={} oracle fp64 {:.1f}s'.format(kind, N, H, W, C, freeze, time.time() - t), flush=True)
    for prec in precisions:
        net.load_state_dict(sd)
        net.to(dev); net.train()
        if freeze:
            net.freeze_batchnorm()
        net.b2_precision = prec
        for m in net.modules():
            if type(m).__name__ == 'B2Dropout':
                m.inject([dm])
        for p in net.parameters():
            p.grad = None
        torch.cuda.synchronize(); t = time.time()
        y = net(x.to(dev))
        y.backward(dy.to(dev))
        torch.cuda.synchronize()
        el = time.time() - t
        yerr = (y.detach().cpu().double() - yo.detach()).abs().max().item() / yo.abs().max().item()
        errs = []
        for k, p in net.named_parameters():
            if not p.requires_grad or sd64[k].grad is None:
                continue
            g = sd64[k].grad
            errs.append(((p.grad.detach().cpu().double() - g).abs().max().item() / (g.abs().max().item() + 1e-30), k))
        errs.sort()
        rs = 0.0
        for k, v in net.state_dict().items():
            if 'running' in k:
                rs = max(rs, (v.cpu().double() - sd64[k].detach()).abs().max().item())
        print('  [{}] {:.2f}s logits relerr {:.3e} | grad relerr median {:.3e} p90 {:.3e} max {:.3e} ({}) | running-stat maxdiff {:.2e}'.format(
            prec, el, yerr, errs[len(errs) // 2][0], errs[int(len(errs) * 0.9)][0], errs[-1][0], errs[-1][1], rs), flush=True)
        net.cpu()


if __name__ == '__main__':
    run('resnet101_deeplab_imagenet', 2, 65, 65, 21, True)
    run('resnet101_deeplab_imagenet', 3, 97, 81, 21, True, precisions=('3xtf32',))
    run('resnet101_deeplabv3plus_imagenet', 3, 64, 64, 19, True)
    run('resnet101_deeplabv3plus_imagenet', 3, 65, 97, 19, True, precisions=('3xtf32',))
    run('resnet101_deeplabv3plus_imagenet', 3, 64, 64, 19, False, precisions=('3xtf32',))
    run('resnet101_deeplab_imagenet', 3, 65, 65, 21, False, precisions=('3xtf32',))

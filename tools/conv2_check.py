"""Quick correctness check of the 2-CTA conv kernel vs the single-CTA kernel and float64 torch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from cutmix_semisup_seg_b200.kernels import ActKernels
from cutmix_semisup_seg_b200.acts import Act
from cutmix_semisup_seg_b200 import lib as L
dev = torch.device('cuda:0')
K = ActKernels(n_split=1)
torch.manual_seed(0)
for (N, H, W, Cin, Cout, k, dil) in [(2, 16, 16, 64, 256, 1, 1), (2, 16, 16, 64, 256, 3, 1), (3, 13, 11, 96, 512, 3, 2), (1, 64, 64, 256, 1024, 1, 1), (2, 24, 24, 128, 304, 3, 12), (1, 5, 5, 64, 256, 3, 1)]:
    pad = dil * (k // 2)
    x = torch.randn(N, Cin, H, W); w = torch.randn(Cout, Cin, k, k) / (Cin * k * k) ** 0.5
    ref = F.conv2d(x.double(), w.double(), padding=pad, dilation=dil)
    xa = Act(x.permute(0, 2, 3, 1).contiguous().to(dev), N, H, W, Cin)
    wd = w.permute(0, 2, 3, 1).contiguous().to(dev)
    res = torch.randn(N, H, W, Cout, device=dev)
    outs = []
    for force1 in (1, 0):
        L.load().b2_debug_set(2, force1)
        out = Act.alloc(N, H, W, Cout, dev)
        K.conv_fwd(xa, wd, Cout, k, k, Cin, Cin, 1, pad, dil, out, addend=Act(res, N, H, W, Cout), relu=True)
        torch.cuda.synchronize()
        outs.append(out.to_nchw().cpu().double())
    L.load().b2_debug_set(2, 0)
    r = torch.relu(ref + res.cpu().permute(0, 3, 1, 2).double())
    e1 = (outs[0] - r).abs().max().item() / r.abs().max().item(); e2 = (outs[1] - r).abs().max().item() / r.abs().max().item()
    print('N{} {}x{} {}->{} k{} d{}: 1cta err {:.2e}  2cta err {:.2e}  1cta-vs-2cta maxdiff {:.2e}'.format(N, H, W, Cin, Cout, k, dil, e1, e2, (outs[0] - outs[1]).abs().max().item()), flush=True)

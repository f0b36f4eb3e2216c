"""One-hot diagnostics for the MN-major wgrad kernel (development tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cutmix_semisup_seg_b200 import ops as O, lib as L
dev = torch.device('cuda:0'); be = O.CudaBackend()
torch.set_printoptions(linewidth=200, precision=3, sci_mode=False)

def run(N, H, W, Cin, Cout, dy, x, variant, n_split=1):
    dw = torch.full((Cout, 1, Cin), -7.0, device=dev)
    L.load().b2_debug_set(1, variant)
    be.conv_wgrad(dy.data_ptr(), N, H, W, Cout, Cout, x.data_ptr(), H, W, Cin, Cin, dw.data_ptr(), O.conv_taps(1, 1, 1, 0), 1,
                  n_split=n_split, device=dev)
    torch.cuda.synchronize()
    L.load().b2_debug_set(1, 0)
    return dw.view(Cout, Cin).cpu()

N, H, W, Cin, Cout = 1, 4, 8, 64, 128
for variant in (0, 1):
    for p0 in (0, 1, 9, 31):
        dy = torch.zeros(N * H * W, Cout); x = torch.zeros(N * H * W, Cin)
        dy[p0] = torch.arange(1, Cout + 1).float(); x[p0] = torch.arange(1, Cin + 1).float() * 0.01
        ref = dy.t() @ x
        got = run(N, H, W, Cin, Cout, dy.to(dev), x.to(dev), variant)
        print('variant', variant, 'p0', p0, 'max|got|', got.abs().max().item(), 'nnz', int((got != 0).sum()), 'n(-7)', int((got == -7).sum()),
              'maxerr', (got - ref).abs().max().item())
        if p0 in (0, 9):
            print(' got[0:3,0:6]', got[0:3, 0:6].tolist()); print(' ref[0:3,0:6]', ref[0:3, 0:6].tolist())
            print(' got[31:34,30:34]', got[31:34, 30:34].tolist()); print(' ref[31:34,30:34]', ref[31:34, 30:34].tolist())
    torch.manual_seed(0)
    dy = torch.randn(N * H * W, Cout); x = torch.randn(N * H * W, Cin)
    ref = dy.t() @ x
    got = run(N, H, W, Cin, Cout, dy.to(dev), x.to(dev), variant)
    print('variant', variant, 'random: maxerr', (got - ref).abs().max().item(), 'max|ref|', ref.abs().max().item(),
          'corr', torch.corrcoef(torch.stack([got.flatten(), ref.flatten()]))[0, 1].item())
    # does got match ref for a sub-block?
    for (r0, r1, c0, c1) in [(0, 32, 0, 32), (32, 64, 0, 32), (0, 32, 32, 64), (64, 128, 0, 64)]:
        print('   block', (r0, r1, c0, c1), 'err', (got[r0:r1, c0:c1] - ref[r0:r1, c0:c1]).abs().max().item())

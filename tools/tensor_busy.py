"""Tensor-pipe occupancy of a CTA-pair convolution launch, from the launch plan and a profiler's cycle count.

    python tools/tensor_busy.py N H W CIN COUT K DIL SM_CYCLES_ELAPSED [SMS]

The kernel's tile walk is deterministic: b2_conv_gemm_plan (host-only, no GPU) returns the pipeline stages every CTA pair executes
(padding-only taps skipped).  A stage = 4 tcgen05.mma M256 x N256 x K8 kind::tf32 instructions = 4 x 128 tensor-pipe cycles on
each SM of the pair, so   busy cycles of an SM = stages of its pair x 512,  and dividing by ncu's sm__cycles_elapsed gives the
fraction of cycles the tensor pipe is busy.  (ncu 2025.2's sm__pipe_tensor_cycles_active_realtime is a sampled counter that
varies 28-66 % between identical back-to-back launches of these kernels: profiles/r02_v18_*.)"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cutmix_semisup_seg_b200 import lib as L


def plan(n, h, w, cin, cout, k, dil, sms=148, n_split=1):
    pad = dil * (k // 2)
    taps = []
    for r in range(k):
        for s in range(k):
            taps += [r * dil - pad, s * dil - pad, r * k + s]
    taps = np.array(taps, dtype=np.int32)
    p = L.ConvParams()
    fake = 0x10000                      # never dereferenced by the planner
    p.a, p.b, p.d = fake, fake, fake
    p.n, p.ih, p.iw, p.k, p.lda = n, h, w, cin, (cin + 3) // 4 * 4
    p.nb, p.tb, p.ldb = cout, k * k, (cin + 3) // 4 * 4
    p.oh, p.ow, p.fh, p.fw, p.ldd = h, w, h, w, (cout + 3) // 4 * 4
    p.ostride, p.ooh, p.oow, p.istride = 1, 0, 0, 1
    p.n_taps, p.taps, p.n_split = k * k, taps.ctypes.data, n_split
    out = (ctypes.c_int64 * 8)()
    L.call('b2_conv_gemm_plan', ctypes.byref(p), int(sms), ctypes.cast(out, ctypes.c_void_p))
    keys = ('bw', 'bh', 'bn', 'tile_pairs', 'stages_total', 'stages_busiest_pair', 'cta_pairs', 'mma_per_stage')
    return dict(zip(keys, list(out)))


if __name__ == '__main__':
    n, h, w, cin, cout, k, dil = [int(v) for v in sys.argv[1:8]]
    cycles = float(sys.argv[8])
    sms = int(sys.argv[9]) if len(sys.argv) > 9 else 148
    pl = plan(n, h, w, cin, cout, k, dil, sms)
    dense = pl['tile_pairs'] * k * k * ((cin + 31) // 32)
    per_stage = pl['mma_per_stage'] * 128
    avg = pl['stages_total'] / pl['cta_pairs'] * per_stage
    worst = pl['stages_busiest_pair'] * per_stage
    print('plan', pl)
    print('stages executed / dense stages: {:.4f}  (padding-only taps skipped)'.format(pl['stages_total'] / dense))
    print('tensor-pipe busy cycles per SM: average {:.0f}, busiest pair {:.0f}; sm__cycles_elapsed {:.0f}'.format(avg, worst, cycles))
    print('tensor-pipe occupancy: {:.1f} % (average SM), {:.1f} % (busiest pair)'.format(100 * avg / cycles, 100 * worst / cycles))

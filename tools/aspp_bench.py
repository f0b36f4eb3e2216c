"""Micro-benchmark of the tensor-core conv kernels on the DeepLab v3+ ASPP shapes (N=16, 2048x64x64 -> 256):
CUDA-event timing per launch (L2 flushed between launches), TFLOP/s vs the measured peak.  Used for the ncu
captures committed under profiles/."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cutmix_semisup_seg_b200.kernels import ActKernels
from cutmix_semisup_seg_b200.acts import Act

dev = torch.device('cuda:0')
K = ActKernels(n_split=1)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)


def timeit(fn, flops, name):
    ts = []
    for i in range(reps + 1):
        flush.fill_(float(i))                      # evict L2 (256 MB write)
        torch.cuda._sleep(600000)                  # let the host run ahead: keeps enqueue latency out of e0..e1
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i > 0:
            ts.append(e0.elapsed_time(e1))
    ms = sum(ts) / len(ts)
    print('{:<44s} {:8.3f} ms  {:8.1f} TFLOP/s'.format(name, ms, flops / ms / 1e9), flush=True)
    return ms


def conv_case(n, h, w, cin, cout, k, dil, name):
    pad = dil * (k // 2)
    x = Act(torch.randn(n, h, w, cin, device=dev), n, h, w, cin)
    wt = torch.randn(cout, k * k, cin, device=dev) * 0.01
    y = Act.alloc(n, h, w, cout, dev)
    g = Act(torch.randn(n, h, w, cout, device=dev), n, h, w, cout)
    dx = Act.alloc(n, h, w, cin, dev)
    dw = torch.zeros(cout, k * k, cin, device=dev)
    wtt, ldb = K.transpose_w(wt, cout, k * k, cin)
    fl = 2.0 * n * h * w * cin * cout * k * k
    timeit(lambda: K.conv_fwd(x, wt, cout, k, k, cin, cin, 1, pad, dil, y), fl, name + ' fprop')
    timeit(lambda: K.conv_dgrad(g, wtt, cin, k, k, cout, ldb, 1, pad, dil, dx), fl, name + ' dgrad')
    timeit(lambda: K.conv_wgrad(g, x, dw, cout, k, k, cin, 1, pad, dil), fl, name + ' wgrad')


if __name__ == '__main__':
    which = sys.argv[2] if len(sys.argv) > 2 else 'all'
    if os.environ.get('B200SEG_TAP_OUTER'):          # A/B of the producer loop order (debug knob 7)
        from cutmix_semisup_seg_b200 import lib as _lib0
        _lib0.load().b2_debug_set(7, int(os.environ['B200SEG_TAP_OUTER']))
        print('debug knob 7 (tap-outer producer order) =', os.environ['B200SEG_TAP_OUTER'])
    if os.environ.get('B200SEG_MAIN_STAGES'):        # A/B of the operand ring depth of the compute-bound CTA-pair kernel (debug knob 10)
        from cutmix_semisup_seg_b200 import lib as _lib1
        _lib1.load().b2_debug_set(10, int(os.environ['B200SEG_MAIN_STAGES']))
        print('debug knob 10 (operand stages of the main build) =', os.environ['B200SEG_MAIN_STAGES'])
    if which in ('all', 'aspp'):
        conv_case(16, 64, 64, 2048, 256, 3, 12, 'ASPP 3x3 d12 2048->256 @64x64 N16')
    if which == 'aspp32':      # the launch shapes of the iteration since the head runs once over both mini-batches (32 images)
        conv_case(32, 64, 64, 2048, 256, 3, 12, 'ASPP 3x3 d12 2048->256 @64x64 N32')
    if which == 'wg':          # weight-gradient A/B: load-balanced plan (knob 14) x operand ring depth (knob 13) on the head / trunk shapes
        from cutmix_semisup_seg_b200 import lib as _lib2
        L2 = _lib2.load()
        for bal in (0, 1):
            for st in (6, 7):
                L2.b2_debug_set(14, bal); L2.b2_debug_set(13, st)
                print('--- balance knob 14 = {}, wgrad ring depth (knob 13) = {}'.format(bal, st), flush=True)
                for (n, h, w, cin, cout, k, dil, name) in ((32, 64, 64, 2048, 256, 3, 12, 'ASPP d12 N32'), (32, 64, 64, 2048, 256, 3, 24, 'ASPP d24 N32'),
                                                           (32, 64, 64, 2048, 256, 3, 36, 'ASPP d36 N32'), (32, 64, 64, 256, 256, 3, 2, 'layer3 3x3 N32'),
                                                           (32, 64, 64, 512, 512, 3, 4, 'layer4 3x3 N32'), (32, 64, 64, 256, 1024, 1, 1, 'layer3 1x1 256->1024 N32')):
                    pad = dil * (k // 2)
                    x = Act(torch.randn(n, h, w, cin, device=dev), n, h, w, cin)
                    g = Act(torch.randn(n, h, w, cout, device=dev), n, h, w, cout)
                    dw = torch.zeros(cout, k * k, cin, device=dev)
                    timeit(lambda: K.conv_wgrad(g, x, dw, cout, k, k, cin, 1, pad, dil), 2.0 * n * h * w * cin * cout * k * k, name + ' wgrad')
                    del x, g, dw
        L2.b2_debug_set(14, 1); L2.b2_debug_set(13, 6)
    if which == 'wgscan':      # scan of (pixel splits, tap rotation) for the ASPP weight gradients (debug knobs 15 / 16)
        from cutmix_semisup_seg_b200 import lib as _lib3
        L3 = _lib3.load()
        for dil in (12, 24, 36):
            n, h, w, cin, cout, k = 32, 64, 64, 2048, 256, 3
            x = Act(torch.randn(n, h, w, cin, device=dev), n, h, w, cin)
            g = Act(torch.randn(n, h, w, cout, device=dev), n, h, w, cout)
            dw = torch.zeros(cout, k * k, cin, device=dev)
            for sp, st in ((1, 0), (2, 0), (2, 3), (3, 3), (3, 1), (4, 0), (4, 2), (4, 1), (6, 1), (8, 0), (8, 1), (8, 2), (9, 1), (9, 4), (16, 0), (16, 1), (16, 5), (32, 1)):
                L3.b2_debug_set(15, sp); L3.b2_debug_set(16, st)
                timeit(lambda: K.conv_wgrad(g, x, dw, cout, k, k, cin, 1, dil, dil), 2.0 * n * h * w * cin * cout * k * k,
                       'ASPP d{} N32 wgrad splits {} step {}'.format(dil, sp, st))
            del x, g, dw
        L3.b2_debug_set(15, 0); L3.b2_debug_set(16, -1)
    if which == 'widepf':      # L2 prefetch of the A stream in 256-channel boxes (debug knob 18) on the K-heavy 1x1 layers
        from cutmix_semisup_seg_b200 import lib as _lib4
        L4 = _lib4.load()
        for (n, h, w, cin, cout, name) in ((32, 64, 64, 1024, 256, 'layer3 1x1 1024->256 N32'), (32, 64, 64, 2048, 512, 'layer4 1x1 2048->512 N32'),
                                           (32, 64, 64, 2048, 256, 'ASPP 1x1 2048->256 N32'), (32, 64, 64, 1024, 2048, 'layer4 ds 1x1 1024->2048 N32'),
                                           (32, 64, 64, 512, 128, 'layer2 1x1 512->128 N32')):
            x = Act(torch.randn(n, h, w, cin, device=dev), n, h, w, cin)
            wt = torch.randn(cout, 1, cin, device=dev) * 0.01
            gate = Act(torch.randn(n, h, w, cout, device=dev), n, h, w, cout)
            sc = torch.rand(cout, device=dev) + 0.5; sh = torch.randn(cout, device=dev)
            outs = []
            for knob in (0, 1, 0, 1):
                L4.b2_debug_set(18, knob)
                y = Act.alloc(n, h, w, cout, dev); y2 = Act.alloc(n, h, w, cout, dev)
                fl = 2.0 * n * h * w * cin * cout
                timeit(lambda: K.conv_fwd(x, wt, cout, 1, 1, cin, cin, 1, 0, 1, y, scale=sc, shift=sh, relu=True), fl, name + ' fwd bn+relu [wide_pf={}]'.format(knob))
                timeit(lambda: K.conv_fwd(x, wt, cout, 1, 1, cin, cin, 1, 0, 1, y2, gate=gate), fl, name + ' dgrad gate [wide_pf={}]'.format(knob))
                outs.append((y.base.clone(), y2.base.clone()))
            print('   bit-identical with / without the prefetch:', all(torch.equal(a, b) for o in outs[1:] for a, b in zip(outs[0], o)), flush=True)
            del x, gate
        L4.b2_debug_set(18, 0)
    if which == 'l3x3':        # the largest compute-bound group of the trunk
        conv_case(32, 64, 64, 256, 256, 3, 2, 'layer3 3x3 d2 256->256 @64x64 N32')
    if which == 'l3':
        conv_case(16, 64, 64, 256, 1024, 1, 1, 'layer3 1x1 256->1024 @64x64 N16')
    if which == 'all':
        conv_case(16, 64, 64, 2048, 256, 3, 36, 'ASPP 3x3 d36 2048->256 @64x64 N16')
        conv_case(16, 64, 64, 2048, 256, 1, 1, 'ASPP 1x1 2048->256 @64x64 N16')
        conv_case(16, 64, 64, 256, 256, 3, 2, 'layer3 3x3 d2 256->256 @64x64 N16')
        conv_case(16, 64, 64, 256, 1024, 1, 1, 'layer3 1x1 256->1024 @64x64 N16')
        conv_case(16, 64, 64, 1024, 256, 1, 1, 'layer3 1x1 1024->256 @64x64 N16')
        conv_case(16, 64, 64, 512, 512, 3, 4, 'layer4 3x3 d4 512->512 @64x64 N16')
        conv_case(16, 128, 128, 304, 256, 3, 1, 'decoder 3x3 304->256 @128x128 N16')
        conv_case(16, 128, 128, 64, 64, 3, 1, 'layer1 3x3 64->64 @128x128 N16')
        conv_case(16, 128, 128, 64, 256, 1, 1, 'layer1 1x1 64->256 @128x128 N16')
    if which == 'epi':
        # timing experiment for the HBM-bound 1x1 layers: where does the per-tile time go?
        # debug knob 3: 1 = epilogue skips its HBM stores, 2 = also skips the TMEM loads; knob 2: force the 1-CTA kernel
        from cutmix_semisup_seg_b200 import lib as _lib
        L = _lib.load()
        for force1 in (0, 1):
            L.b2_debug_set(2, force1)
            for dbg in (0, 1, 2):
                L.b2_debug_set(3, dbg)
                tag = ' [1cta={} dbg={}]'.format(force1, dbg)
                n, h, w = 16, 64, 64
                for cin, cout in ((256, 1024), (1024, 256)):
                    x = Act(torch.randn(n, h, w, cin, device=dev), n, h, w, cin)
                    wt = torch.randn(cout, 1, cin, device=dev) * 0.01
                    y = Act.alloc(n, h, w, cout, dev)
                    fl = 2.0 * n * h * w * cin * cout
                    timeit(lambda: K.conv_fwd(x, wt, cout, 1, 1, cin, cin, 1, 0, 1, y), fl, '1x1 {}->{}'.format(cin, cout) + tag)
        L.b2_debug_set(3, 0); L.b2_debug_set(2, 0)
    if which == 'trace':
        # pipeline timeline of CTA 0 of the 2-CTA kernel (clock64 stamps, see b2_debug_trace in include/b200seg.h)
        from cutmix_semisup_seg_b200 import lib as _lib
        L = _lib.load()
        n, h, w = 16, 64, 64
        for cin, cout, dbg in ((256, 1024, 0), (256, 1024, 2), (1024, 256, 0), (1024, 256, 2)):
            x = Act(torch.randn(n, h, w, cin, device=dev), n, h, w, cin)
            wt = torch.randn(cout, 1, cin, device=dev) * 0.01
            y = Act.alloc(n, h, w, cout, dev)
            buf = torch.zeros(2048, dtype=torch.int64, device=dev)
            K.conv_fwd(x, wt, cout, 1, 1, cin, cin, 1, 0, 1, y)
            flush.fill_(1.0); torch.cuda._sleep(600000)
            L.b2_debug_set(3, dbg); L.b2_debug_trace(buf.data_ptr())
            K.conv_fwd(x, wt, cout, 1, 1, cin, cin, 1, 0, 1, y)
            torch.cuda.synchronize()
            L.b2_debug_trace(None); L.b2_debug_set(3, 0)
            b = buf.cpu().numpy()
            t0 = min(v for v in b if v > 0)
            kb = cin // 32
            prod = [v - t0 for v in b[0:512] if v > 0]
            mma = [v - t0 for v in b[512:1024] if v > 0]
            tile = [v - t0 for v in b[1024:1536] if v > 0]
            epi = [v - t0 for v in b[1536:2048] if v > 0]
            print('--- 1x1 {}->{} dbg={} kblocks/tile={}'.format(cin, cout, dbg, kb))
            print('producer stage issue (first 48):', prod[:48])
            print('mma stage acquired   (first 48):', mma[:48])
            print('mma tile [begin, acc acquired]*:', tile[:24])
            print('epi  tile [wait begin, tfull ]*:', epi[:24])
            print('last stamps: prod {} mma {} tile {} epi {}'.format(prod[-1] if prod else 0, mma[-1] if mma else 0, tile[-1] if tile else 0, epi[-1] if epi else 0))
    if which == 'trace3':
        # pipeline timeline of CTA 0 for the compute-bound 3x3 shapes: per-stage cadence of the MMA issuer (cycles between
        # consecutive `full` acquisitions), of the producer, and the per-tile waits (accumulator hand-over, epilogue)
        import numpy as np
        from cutmix_semisup_seg_b200 import lib as _lib
        L = _lib.load()
        for (n, h, w, cin, cout, k, dil, name) in ((32, 64, 64, 256, 256, 3, 2, 'layer3 3x3 d2 256->256 N32'),
                                                   (16, 64, 64, 2048, 256, 3, 12, 'ASPP 3x3 d12 2048->256 N16'),
                                                   (32, 64, 64, 512, 512, 3, 4, 'layer4 3x3 d4 512->512 N32')):
            pad = dil * (k // 2)
            x = Act(torch.randn(n, h, w, cin, device=dev), n, h, w, cin)
            wt = torch.randn(cout, k * k, cin, device=dev) * 0.01
            y = Act.alloc(n, h, w, cout, dev)
            sc = torch.rand(cout, device=dev) + 0.5; sh = torch.randn(cout, device=dev)
            buf = torch.zeros(2048, dtype=torch.int64, device=dev)
            run = lambda: K.conv_fwd(x, wt, cout, k, k, cin, cin, 1, pad, dil, y, scale=sc, shift=sh, relu=True)
            run(); flush.fill_(1.0); torch.cuda._sleep(600000)
            L.b2_debug_trace(buf.data_ptr()); run(); torch.cuda.synchronize(); L.b2_debug_trace(None)
            b = buf.cpu().numpy()
            t0 = min(v for v in b if v > 0)
            prod = np.array([v - t0 for v in b[0:512] if v > 0]); mma = np.array([v - t0 for v in b[512:1024] if v > 0])
            tile = np.array([v - t0 for v in b[1024:1536] if v > 0]); epi = np.array([v - t0 for v in b[1536:2048] if v > 0])
            stages_per_tile = 9 * (cin // 32)
            dm, dp = np.diff(mma), np.diff(prod)
            print('--- {}: {} stages per tile (ideal 512 cycles per stage at the tf32 rate)'.format(name, stages_per_tile))
            print('  MMA stage cadence  : median {:.0f}  mean {:.0f}  p90 {:.0f}  max {:.0f} cycles  (n={})'.format(np.median(dm), dm.mean(), np.percentile(dm, 90), dm.max(), len(dm)))
            print('  producer cadence   : median {:.0f}  mean {:.0f}  p90 {:.0f}  max {:.0f} cycles'.format(np.median(dp), dp.mean(), np.percentile(dp, 90), dp.max()))
            big = np.argsort(dm)[-6:][::-1]
            print('  largest MMA gaps at stage index (cycles):', [(int(i), int(dm[i])) for i in big])
            tl = tile.reshape(-1, 2) if len(tile) % 2 == 0 else tile[:-1].reshape(-1, 2)
            print('  tiles of CTA 0: begin ->', [int(v) for v in tl[:8, 0]], ' wait for the accumulator (cycles):', [int(v) for v in (tl[:, 1] - tl[:, 0])[:8]])
            ep = epi.reshape(-1, 2) if len(epi) % 2 == 0 else epi[:-1].reshape(-1, 2)
            print('  epilogue warp 4: wait begin ->', [int(v) for v in ep[:8, 0]], ' until the accumulator is full:', [int(v) for v in (ep[:, 1] - ep[:, 0])[:8]])
            if len(tl) > 1:
                print('  tile period (MMA warp, begin to begin):', [int(v) for v in np.diff(tl[:, 0])[:8]], ' = {:.0f} cycles per stage'.format(np.diff(tl[:, 0]).mean() / stages_per_tile))
    if which == 'l3full':
        # the two in-situ flavours of the HBM-bound layer3 1x1 (256 -> 1024 channels): forward with folded BN +
        # residual + ReLU, and the dgrad with partial-gradient addend + ReLU gate + fused BN statistics (+ residual sub)
        from cutmix_semisup_seg_b200 import lib as _lib
        L = _lib.load()
        n, h, w, cin, cout = 16, 64, 64, 256, 1024
        x = Act(torch.randn(n, h, w, cin, device=dev), n, h, w, cin)
        wt = torch.randn(cout, 1, cin, device=dev) * 0.01
        y = Act.alloc(n, h, w, cout, dev)
        res = Act(torch.randn(n, h, w, cout, device=dev), n, h, w, cout)
        gate = Act(torch.randn(n, h, w, cout, device=dev), n, h, w, cout)
        sub = Act(torch.randn(n, h, w, cout, device=dev), n, h, w, cout)
        sc = torch.rand(cout, device=dev) + 0.5; sh = torch.randn(cout, device=dev)
        fl = 2.0 * n * h * w * cin * cout
        g = Act(torch.randn(n, h, w, cin, device=dev), n, h, w, cin)
        wtt, ldb = K.transpose_w(torch.randn(cin, 1, cout, device=dev) * 0.01, cin, 1, cout)
        for dbg in (0, 4):
            L.b2_debug_set(3, dbg)
            tag = ' [prefetch={}]'.format(dbg >> 2)
            timeit(lambda: K.conv_fwd(x, wt, cout, 1, 1, cin, cin, 1, 0, 1, y), fl, 'fwd plain' + tag)
            timeit(lambda: K.conv_fwd(x, wt, cout, 1, 1, cin, cin, 1, 0, 1, y, scale=sc, shift=sh, addend=res, relu=True), fl, 'fwd bn+residual+relu' + tag)
            timeit(lambda: K.conv_dgrad(g, wtt, cout, 1, 1, cin, ldb, 1, 0, 1, y, addend=res, gate=gate), fl, 'dgrad addend+gate' + tag)
            timeit(lambda: K.conv_dgrad(g, wtt, cout, 1, 1, cin, ldb, 1, 0, 1, y, addend=res, gate=gate, want_stats=True, stats_sub=sub), fl, 'dgrad addend+gate+stats+sub' + tag)
        L.b2_debug_set(3, 0)
    if which == 'pf':
        # asynchronous epilogue-operand prefetch (cp.async slots, conv_epilogue.cuh) on / off: timing of the HBM-bound
        # flavours and a bit-exactness check of the two builds of the 2-CTA kernel (debug knob 4 = K*taps threshold)
        from cutmix_semisup_seg_b200 import lib as _lib, ops as O
        L = _lib.load()
        n, h, w = 16, 64, 64

        def flavours(cin, cout, k, dil, tag):
            pad = dil * (k // 2)
            x = Act(torch.randn(n, h, w, cin, device=dev), n, h, w, cin)
            wt = torch.randn(cout, k * k, cin, device=dev) * 0.01
            res = Act(torch.randn(n, h, w, cout, device=dev), n, h, w, cout)
            gate = Act(torch.randn(n, h, w, cout, device=dev), n, h, w, cout)
            sc = torch.rand(cout, device=dev) + 0.5; sh = torch.randn(cout, device=dev)
            fl = 2.0 * n * h * w * cin * cout * k * k
            outs = {}
            for thr in (0, 1 << 20):
                L.b2_debug_set(4, thr)
                t = ' {} [pf={}]'.format(tag, int(thr > 0))
                y1 = Act.alloc(n, h, w, cout, dev); y2 = Act.alloc(n, h, w, cout, dev); y3 = Act.alloc(n, h, w, cout, dev)
                timeit(lambda: K.conv_fwd(x, wt, cout, k, k, cin, cin, 1, pad, dil, y1, scale=sc, shift=sh, addend=res, relu=True), fl, 'fwd bn+residual+relu' + t)
                timeit(lambda: K.conv_fwd(x, wt, cout, k, k, cin, cin, 1, pad, dil, y2, addend=res, gate=gate), fl, 'gemm addend+gate' + t)
                stats = [None]

                def run3b():
                    stats[0] = K.be.conv_gemm(x.ptr, x.n, x.h, x.w, cin, x.ld, wt.data_ptr(), cout, k * k, cin, y3.ptr, h, w, h, w,
                                              y3.ld, O.conv_taps(k, k, dil, pad), addend=res.ptr, ld_add=res.ld, gate=gate.ptr,
                                              ld_gate=gate.ld, want_stats=True, device=dev)
                timeit(run3b, fl, 'gemm addend+gate+stats' + t)
                outs[thr] = (y1.base.clone(), y2.base.clone(), y3.base.clone(), stats[0][0].clone())
            same = all(torch.equal(a, b) for a, b in zip(outs[0], outs[1 << 20]))
            print('   bit-identical with / without prefetch: {}'.format(same), flush=True)
            L.b2_debug_set(4, 0)

        flavours(256, 1024, 1, 1, '1x1 256->1024')
        flavours(512, 2048, 1, 1, '1x1 512->2048')
        flavours(1024, 256, 1, 1, '1x1 1024->256')
        flavours(128, 512, 1, 1, '1x1 128->512')
        flavours(256, 256, 3, 2, '3x3 d2 256->256')
        # ragged case: odd spatial size, channel tail (nb % 32 != 0), batch 3
        n, h, w = 3, 33, 41
        flavours(96, 304, 1, 1, 'ragged 1x1 96->304')
    if which == 'tma':
        # TMA epilogue (conv_epilogue_tma.cuh, debug knob 8) vs the register epilogue on the HBM-bound 1x1 layers of the
        # batched trunk (N = 32 at 64 x 64): timing of every flavour the step uses and a bit-exactness check of outputs / fused
        # statistics (the two epilogues perform the same fp32 operations in the same order)
        from cutmix_semisup_seg_b200 import lib as _lib, ops as O
        L = _lib.load()

        def flavours(n, h, w, cin, cout, tag, time_it=True):
            x = Act(torch.randn(n, h, w, cin, device=dev), n, h, w, cin)
            wt = torch.randn(cout, 1, cin, device=dev) * 0.05
            res = Act(torch.randn(n, h, w, cout, device=dev), n, h, w, cout)
            gate = Act(torch.randn(n, h, w, cout, device=dev), n, h, w, cout)
            sc = torch.rand(cout, device=dev) + 0.5; sh = torch.randn(cout, device=dev)
            fl = 2.0 * n * h * w * cin * cout
            outs = {}
            for knob in (0, 1):
                L.b2_debug_set(8, knob)
                t = ' {} [tma={}]'.format(tag, knob)
                ys = [Act.alloc(n, h, w, cout, dev) for _ in range(6)]
                for y in ys:
                    y.base.fill_(0.5)
                stats = [None]

                def plain(): K.conv_fwd(x, wt, cout, 1, 1, cin, cin, 1, 0, 1, ys[0])
                def bnres(): K.conv_fwd(x, wt, cout, 1, 1, cin, cin, 1, 0, 1, ys[1], scale=sc, shift=sh, addend=res, relu=True)
                def addgate(): K.conv_fwd(x, wt, cout, 1, 1, cin, cin, 1, 0, 1, ys[2], addend=res, gate=gate)
                def gateonly(): K.conv_fwd(x, wt, cout, 1, 1, cin, cin, 1, 0, 1, ys[5], gate=gate)

                def addgatestats():
                    stats[0] = K.be.conv_gemm(x.ptr, 1, 1, x.rows, cin, x.ld, wt.data_ptr(), cout, 1, cin, ys[3].ptr, 1, x.rows, 1, x.rows,
                                              ys[3].ld, O.conv_taps(1, 1, 1, 0), addend=res.ptr, ld_add=res.ld, gate=gate.ptr,
                                              ld_gate=gate.ld, want_stats=True, device=dev)

                def accum():
                    K.be.conv_gemm(x.ptr, 1, 1, x.rows, cin, x.ld, wt.data_ptr(), cout, 1, cin, ys[4].ptr, 1, x.rows, 1, x.rows,
                                   ys[4].ld, O.conv_taps(1, 1, 1, 0), accumulate=True)
                for fn, name in ((plain, 'fwd plain'), (bnres, 'fwd bn+residual+relu'), (addgate, 'dgrad addend+gate'),
                                 (addgatestats, 'dgrad addend+gate+stats'), (gateonly, 'dgrad gate only')):
                    if time_it:
                        timeit(fn, fl, name + t)
                    else:
                        fn()
                ys[4].base.fill_(0.5); accum(); torch.cuda.synchronize()
                outs[knob] = [y.base.clone() for y in ys] + [stats[0][0].clone()]
            L.b2_debug_set(8, 1)
            same = [bool(torch.equal(a, b)) for a, b in zip(outs[0], outs[1])]
            worst = max(float((a - b).abs().max()) for a, b in zip(outs[0], outs[1]))
            print('   {}: bit-identical register / TMA epilogue (plain, bn+res, add+gate, stats out, accumulate, gate only, stats): {}  max abs diff {:.3e}'
                  .format(tag, same, worst), flush=True)

        flavours(32, 64, 64, 256, 1024, '1x1 256->1024 N32')
        flavours(32, 64, 64, 512, 2048, '1x1 512->2048 N32')
        flavours(32, 128, 128, 64, 256, '1x1 64->256 N32 @128')
        flavours(32, 64, 64, 128, 512, '1x1 128->512 N32')
        # geometry cases, checked only: 3-D boxes through the generic conv path are covered by the pytest suite; here ragged
        # pixel counts / channel tails of the flattened form
        flavours(3, 33, 41, 96, 304, 'ragged 1x1 96->304 (4059 px)', time_it=False)
        flavours(1, 1, 200, 256, 1000, 'ragged 1x1 256->1000 (200 px)', time_it=False)
    if which == 'iso':
        # pipeline isolation (timing experiments, results are garbage): conv knob 3 / wgrad knob 6, bit 8 = producer
        # skips the TMA loads, bit 16 = issuer skips the MMAs.  full ~ max(parts): bound by that part; full ~ sum: latency.
        from cutmix_semisup_seg_b200 import lib as _lib
        L = _lib.load()
        for (n, h, w, cin, cout, k, dil, name) in ((16, 64, 64, 256, 256, 3, 2, 'layer3 3x3 d2 256->256'),
                                                   (16, 64, 64, 1024, 256, 1, 1, 'layer3 1x1 1024->256'),
                                                   (16, 64, 64, 2048, 256, 3, 12, 'ASPP 3x3 d12 2048->256')):
            pad = dil * (k // 2)
            x = Act(torch.randn(n, h, w, cin, device=dev), n, h, w, cin)
            g = Act(torch.randn(n, h, w, cout, device=dev), n, h, w, cout)
            wt = torch.randn(cout, k * k, cin, device=dev) * 0.01
            y = Act.alloc(n, h, w, cout, dev)
            dw = torch.zeros(cout, k * k, cin, device=dev)
            fl = 2.0 * n * h * w * cin * cout * k * k
            for dbg in (0, 8, 16, 24):
                L.b2_debug_set(3, dbg)
                timeit(lambda: K.conv_fwd(x, wt, cout, k, k, cin, cin, 1, pad, dil, y), fl, name + ' fprop [dbg={}]'.format(dbg))
            L.b2_debug_set(3, 0)
            for force1 in (1, 0):
                L.b2_debug_set(5, force1)
                for dbg in (0, 8, 16, 24):
                    L.b2_debug_set(6, dbg)
                    timeit(lambda: K.conv_wgrad(g, x, dw, cout, k, k, cin, 1, pad, dil), fl, name + ' wgrad [2cta={} dbg={}]'.format(1 - force1, dbg))
            L.b2_debug_set(6, 0); L.b2_debug_set(5, 0)
    if which == 'major':
        # which MN-major operand costs the tensor-pipe rate?  MMA-only runs (wgrad knob 6 bit 8: no TMA traffic) with the
        # A / B descriptors flipped to K-major SWIZZLE_128B (bits 32 / 64; timing only, the products are garbage)
        from cutmix_semisup_seg_b200 import lib as _lib
        L = _lib.load()
        n, h, w, cin, cout, k, dil = 16, 64, 64, 256, 256, 3, 2
        x = Act(torch.randn(n, h, w, cin, device=dev), n, h, w, cin)
        g = Act(torch.randn(n, h, w, cout, device=dev), n, h, w, cout)
        dw = torch.zeros(cout, k * k, cin, device=dev)
        fl = 2.0 * n * h * w * cin * cout * k * k
        for force1 in (1, 0):
            L.b2_debug_set(5, force1)
            for bits, tag in ((0, 'A mn, B mn (product)'), (32, 'A k , B mn'), (64, 'A mn, B k '), (96, 'A k , B k ')):
                L.b2_debug_set(6, 8 | bits)
                timeit(lambda: K.conv_wgrad(g, x, dw, cout, k, k, cin, 1, 2, dil), fl, '3x3 d2 256->256 wgrad MMA-only [2cta={}] {}'.format(1 - force1, tag))
        L.b2_debug_set(6, 0); L.b2_debug_set(5, 0)
    if which == 'wg2':
        # weight-gradient kernel: single-CTA vs CTA-pair build (debug knob 5) on the hot-path shapes, with a bit-level
        # comparison of the two results (same K order per accumulator => expected identical up to the split count)
        from cutmix_semisup_seg_b200 import lib as _lib
        L = _lib.load()
        for (n, h, w, cin, cout, k, dil, name) in ((16, 64, 64, 256, 256, 3, 2, 'layer3 3x3 d2 256->256'),
                                                   (16, 64, 64, 256, 1024, 1, 1, 'layer3 1x1 256->1024'),
                                                   (16, 64, 64, 1024, 256, 1, 1, 'layer3 1x1 1024->256'),
                                                   (16, 64, 64, 512, 512, 3, 4, 'layer4 3x3 d4 512->512'),
                                                   (16, 64, 64, 2048, 256, 3, 12, 'ASPP 3x3 d12 2048->256'),
                                                   (16, 64, 64, 512, 2048, 1, 1, 'layer4 1x1 512->2048'),
                                                   (16, 128, 128, 256, 256, 3, 1, 'decoder 3x3 256->256'),
                                                   (16, 128, 128, 304, 256, 3, 1, 'decoder 3x3 304->256')):
            pad = dil * (k // 2)
            x = Act(torch.randn(n, h, w, cin, device=dev), n, h, w, cin)
            g = Act(torch.randn(n, h, w, cout, device=dev), n, h, w, cout)
            fl = 2.0 * n * h * w * cin * cout * k * k
            res = []
            for force1 in (1, 0):
                L.b2_debug_set(5, force1)
                dw = torch.zeros(cout, k * k, cin, device=dev)
                timeit(lambda: K.conv_wgrad(g, x, dw, cout, k, k, cin, 1, pad, dil), fl, name + ' wgrad [2cta={}]'.format(1 - force1))
                res.append(dw.clone())
            d = (res[0] - res[1]).abs().max().item() / res[0].abs().max().item()
            print('   max rel diff 1-CTA vs 2-CTA: {:.3e}'.format(d), flush=True)
        L.b2_debug_set(5, 0)

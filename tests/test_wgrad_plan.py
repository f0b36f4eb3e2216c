"""CPU: the weight-gradient kernel's launch plan, checked through the C ABI without a GPU (b2_conv_wgrad_plan_check runs the
product's own planner and box-walk code on the host).  The TMA producer walks the pixel boxes of a work unit, the MMA issuer
counts them in closed form: both must agree for every unit, or one pipeline role would wait forever on the device."""
import ctypes

import numpy as np
import pytest

from cutmix_semisup_seg_b200 import lib as L


def _taps(k, dil, pad):
    t = []
    for r in range(k):
        for s in range(k):
            t += [r * dil - pad, s * dil - pad, r * k + s]
    return np.array(t, dtype=np.int32)


def _check(n, oh, ow, m, c, k, dil, pad, istride=1, kchunk=0, n_split=1, max_ctas=0, ih=None, iw=None):
    taps = _taps(k, dil, pad)
    p = L.WgradParams()
    fake = 0x10000                      # never dereferenced by the planner
    p.dy, p.x, p.dw = fake, fake, fake
    if n_split > 1:
        p.dy_lo, p.x_lo = fake, fake
    p.n, p.oh, p.ow, p.m, p.ldy = n, oh, ow, m, (m + 3) // 4 * 4
    p.ih = ih if ih is not None else (oh - 1) * istride + 1
    p.iw = iw if iw is not None else (ow - 1) * istride + 1
    p.c, p.ldx, p.istride = c, (c + 3) // 4 * 4, istride
    p.n_taps, p.taps, p.tw = k * k, taps.ctypes.data, k * k
    p.accumulate, p.n_split, p.max_ctas, p.kchunk = 0, n_split, max_ctas, kchunk
    out = (ctypes.c_int64 * 5)()
    L.call('b2_conv_wgrad_plan_check', ctypes.byref(p), ctypes.cast(out, ctypes.c_void_p))
    return list(out)


HOT_PATH = [
    # (n, oh, ow, m, c, k, dil, pad)   cfg3 / cfg2 shapes of SURVEY.md appendix A
    (16, 64, 64, 256, 2048, 3, 12, 12), (16, 64, 64, 256, 2048, 3, 24, 24), (16, 64, 64, 256, 2048, 3, 36, 36),
    (32, 64, 64, 256, 256, 3, 2, 2), (32, 64, 64, 512, 512, 3, 4, 4), (16, 128, 128, 256, 304, 3, 1, 1),
    (1, 1, 131072, 1024, 256, 1, 1, 0), (1, 1, 2097152, 64, 160, 1, 1, 0), (32, 128, 128, 64, 64, 3, 1, 1),
    (16, 41, 41, 21, 2048, 3, 12, 12), (10, 41, 41, 256, 256, 3, 2, 2), (2, 9, 9, 48, 256, 1, 1, 0),
]


# weight-gradient shapes of the architectures added after the last GPU run (DeepLab v3 head, ResNet U-Net decoder, DenseNet-161
# U-Net at 224x224 / batch 16 or 32 in the batched trunk): Cin = 96 + 48 k is not a multiple of 32, Cout = 48 / 192 / 1056 / 2208
NEW_ARCH = [
    (16, 64, 64, 256, 256, 3, 1, 1), (16, 64, 64, 21, 256, 1, 1, 0),                                   # DeepLab v3 head
    (16, 16, 16, 1024, 2048, 1, 1, 0), (16, 32, 32, 512, 1024, 3, 1, 1), (16, 64, 64, 256, 512, 3, 1, 1),      # ResNet U-Net
    (16, 128, 128, 64, 256, 3, 1, 1), (16, 256, 256, 64, 64, 3, 1, 1), (16, 512, 512, 64, 64, 3, 1, 1), (16, 512, 512, 11, 64, 1, 1, 0),
    (1, 1, 32 * 56 * 56, 192, 96, 1, 1, 0), (1, 1, 32 * 56 * 56, 192, 336, 1, 1, 0), (32, 56, 56, 48, 192, 3, 1, 1),       # dense block 1
    (1, 1, 32 * 56 * 56, 192, 384, 1, 1, 0), (1, 1, 32 * 28 * 28, 192, 720, 1, 1, 0), (32, 28, 28, 48, 192, 3, 1, 1),      # transition 1, block 2
    (1, 1, 32 * 14 * 14, 192, 2064, 1, 1, 0), (1, 1, 32 * 14 * 14, 1056, 2112, 1, 1, 0), (32, 14, 14, 48, 192, 3, 1, 1),   # block 3, transition 3
    (1, 1, 32 * 7 * 7, 192, 2160, 1, 1, 0), (32, 7, 7, 48, 192, 3, 1, 1),                              # block 4
    (1, 1, 16 * 14 * 14, 2208, 2112, 1, 1, 0), (16, 14, 14, 768, 2208, 3, 1, 1), (16, 28, 28, 384, 768, 3, 1, 1),          # line0, decoder
    (16, 56, 56, 96, 384, 3, 1, 1), (16, 112, 112, 96, 96, 3, 1, 1), (16, 224, 224, 64, 96, 3, 1, 1), (16, 224, 224, 2, 64, 1, 1, 0),
    (1, 1, 32 * 112 * 112, 96, 160, 1, 1, 0),                                                          # DenseNet stem (im2col, K padded to 160)
]


@pytest.mark.parametrize('shape', NEW_ARCH)
def test_new_architecture_shapes_have_consistent_stage_counts(shape):
    for n_split, kchunk in ((1, 0), (3, 1024)):                      # throughput mode and the 3xTF32 parity mode
        splits, units, stages, bad, pair = _check(*shape, n_split=n_split, kchunk=kchunk)
        assert bad == 0 and units >= 1 and splits >= 1 and stages >= units, (shape, n_split)


@pytest.mark.parametrize('shape', HOT_PATH)
def test_hot_path_shapes_have_consistent_stage_counts(shape):
    splits, units, stages, bad, pair = _check(*shape)
    assert bad == 0 and units >= 1 and splits >= 1 and stages >= units


def test_cta_pair_kernel_selection_rule():
    assert _check(16, 64, 64, 256, 2048, 3, 12, 12)[4] == 1          # M % 256 == 0 and C % 64 == 0
    assert _check(16, 128, 128, 256, 304, 3, 1, 1)[4] == 0           # C = 304: single-CTA kernel, balanced N tiles
    assert _check(16, 41, 41, 21, 2048, 3, 12, 12)[4] == 0           # 21 classes
    assert _check(16, 64, 64, 256, 2048, 3, 12, 12, max_ctas=1)[4] == 0


def test_padding_taps_are_skipped_but_every_unit_runs_a_stage():
    # dilation 36 on a 20x20 map: most taps read only padding for most boxes; a unit with no contributing box still
    # runs one (zero) stage so that its accumulator is written
    splits, units, stages, bad, _ = _check(2, 20, 20, 256, 256, 3, 36, 36)
    dense = _check(2, 20, 20, 256, 256, 3, 1, 1)
    assert bad == 0 and dense[3] == 0
    assert units <= stages < dense[2]


def test_random_geometries_agree():
    rs = np.random.RandomState(0)
    for _ in range(300):
        k = int(rs.choice([1, 3]))
        dil = int(rs.randint(1, 40)) if k == 3 else 1
        pad = dil * (k // 2) if rs.rand() < 0.7 else int(rs.randint(0, 20)) * (k // 2)
        s = int(rs.choice([1, 1, 2]))
        oh, ow, n = int(rs.randint(1, 70)), int(rs.randint(1, 70)), int(rs.randint(1, 6))
        m, c = int(rs.choice([19, 64, 256, 512])), int(rs.choice([48, 64, 160, 256, 304]))
        ih = (oh - 1) * s + dil * (k - 1) + 1 - 2 * pad
        iw = (ow - 1) * s + dil * (k - 1) + 1 - 2 * pad
        if ih < 1 or iw < 1:
            continue
        out = _check(n, oh, ow, m, c, k, dil, pad, istride=s, kchunk=int(rs.choice([0, 0, 256])),
                     max_ctas=int(rs.choice([0, 0, 7, 40])), ih=ih, iw=iw)
        assert out[3] == 0, (n, oh, ow, m, c, k, dil, pad, s, out)


def test_bad_arguments_are_reported_not_crashed():
    with pytest.raises(L.B2Error):
        _check(0, 8, 8, 64, 64, 3, 1, 1)


def _balance(n, oh, ow, m, c, k, dil, pad, knob=1):
    lib = L.load()
    lib.b2_debug_set(14, knob)
    try:
        taps = _taps(k, dil, pad)
        p = L.WgradParams()
        fake = 0x10000
        p.dy, p.x, p.dw = fake, fake, fake
        p.n, p.oh, p.ow, p.m, p.ldy = n, oh, ow, m, (m + 3) // 4 * 4
        p.ih, p.iw = oh, ow
        p.c, p.ldx, p.istride = c, (c + 3) // 4 * 4, 1
        p.n_taps, p.taps, p.tw = k * k, taps.ctypes.data, k * k
        p.accumulate, p.n_split, p.max_ctas, p.kchunk = 0, 1, 0, 0
        out = (ctypes.c_int64 * 5)()
        L.call('b2_conv_wgrad_plan_balance', ctypes.byref(p), ctypes.cast(out, ctypes.c_void_p))
        chk = (ctypes.c_int64 * 5)()
        L.call('b2_conv_wgrad_plan_check', ctypes.byref(p), ctypes.cast(chk, ctypes.c_void_p))
        assert chk[3] == 0 and chk[2] == out[3] and chk[0] == out[0]      # rotated units: producer walk == closed form, same totals
        return dict(splits=out[0], tap_step=out[1], worst=out[2], total=out[3], workers=out[4])
    finally:
        lib.b2_debug_set(14, 1)


@pytest.mark.parametrize('dil,min_gain', [(12, 1.08), (24, 1.2), (36, 1.8)])
def test_dilated_weight_gradient_units_are_load_balanced(dil, min_gain):
    """ASPP 3x3 weight gradients of the batched-head iteration (32 images, 64 x 64 map, 2048 -> 256): one unit per CTA pair made the
    launch as long as the centre tap although the off-centre taps skip their padding boxes; the balanced plan (more pixel splits,
    taps rotated from split to split) brings the busiest CTA pair close to the mean.  Total work is unchanged."""
    old = _balance(32, 64, 64, 256, 2048, 3, dil, dil, knob=0)
    new = _balance(32, 64, 64, 256, 2048, 3, dil, dil, knob=1)
    assert old['splits'] == 1 and old['tap_step'] == 0 and old['worst'] == 32 * 64 * 64 // 32     # the centre tap: every box
    assert new['total'] == old['total']
    assert new['splits'] > 1 and old['worst'] / new['worst'] >= min_gain
    assert new['worst'] * new['workers'] <= 1.25 * new['total']


def test_undilated_weight_gradient_plans_are_unchanged():
    for case in [(32, 64, 64, 256, 256, 3, 1, 1), (32, 64, 64, 1024, 256, 1, 1, 0), (32, 128, 128, 64, 64, 3, 1, 1)]:
        assert _balance(*case, knob=1) == _balance(*case, knob=0)


def test_conv_gemm_plan_of_the_aspp_launch():
    """b2_conv_gemm_plan (host only): tile walk of the CTA-pair kernel on the ASPP 3x3 d12 launch of the batched-head iteration
    (32 images, 64 x 64 map, 2048 -> 256) -- 512 tile pairs on 74 CTA pairs, 64 K blocks per tap, and only the top / bottom tile
    rows skip the three taps that lie in the padding (tools/tensor_busy.py turns this into the tensor-pipe occupancy)."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))
    import tensor_busy
    pl = tensor_busy.plan(32, 64, 64, 2048, 256, 3, 12)
    assert pl['bw'] * pl['bh'] * pl['bn'] == 128 and pl['tile_pairs'] == 512 and pl['cta_pairs'] == 74 and pl['mma_per_stage'] == 4
    dense = 512 * 9 * 64
    assert 0.85 * dense < pl['stages_total'] < dense
    assert pl['stages_busiest_pair'] <= 7 * 9 * 64 and pl['stages_busiest_pair'] * 74 >= pl['stages_total']
    # dilation 36 on a 64 x 64 map: most off-centre taps of a tile lie in the padding
    assert tensor_busy.plan(32, 64, 64, 2048, 256, 3, 36)['stages_total'] < 0.75 * dense
    # a 1 x 1 layer has nothing to skip
    pl1 = tensor_busy.plan(32, 64, 64, 1024, 256, 1, 1)
    assert pl1['stages_total'] == 512 * 32


def _choose_box(ow, oh, n, max_rows=128, istride=1):
    """Python statement of b2_choose_box (conv_gemm.cu): the (bw, bh, bn) pixel box of an M tile."""
    best, pick = -1.0, (1, 1, 1)
    max_b = 256 // istride
    for bw in range(1, min(ow, max_rows, max_b) + 1):
        for bh in range(1, min(oh, max_b) + 1):
            if bw * bh > max_rows:
                break
            bn = 1
            if bw == ow and bh == oh:
                bn = max(1, min(n, max_rows // (bw * bh)))
            tiles = -(-ow // bw) * -(-oh // bh) * -(-n // bn)
            score = (ow * oh * n) / (tiles * max_rows) + 1e-6 * bw
            if score > best:
                best, pick = score, (bw, bh, bn)
    return pick


@pytest.mark.parametrize('case', [
    # n, h, w, cin, cout, k, dil  -- tile counts that are 1, powers of two, odd, prime (the fast-division multipliers), dilations
    (32, 64, 64, 2048, 256, 3, 12), (32, 64, 64, 2048, 256, 3, 36), (3, 33, 41, 96, 304, 3, 2), (5, 7, 300, 64, 512, 3, 5),
    (1, 129, 3, 32, 1024, 1, 1), (7, 19, 23, 256, 768, 3, 7), (2, 128, 128, 304, 256, 3, 1), (10, 41, 41, 2048, 512, 3, 24),
], ids=lambda c: 'x'.join(map(str, c)))
def test_conv_gemm_plan_matches_a_python_walk_of_the_tiles(case):
    """b2_conv_gemm_plan runs the kernel's own decode_tile on the host: division by multiplication with precomputed reciprocals
    (tc::FastDiv) and tap masks from the row / column tables.  An independent walk with Python integers must give the same box,
    tile-pair count and stage totals (debug knob 17 = 0: the same through the tap loop instead of the tables)."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))
    import tensor_busy
    n, h, w, cin, cout, k, dil = case
    pad = dil * (k // 2)
    bw, bh, bn = _choose_box(w, h, n)
    tw, th, tn = -(-w // bw), -(-h // bh), -(-n // bn)
    m_tiles = tw * th * tn
    n_tiles_n = -(-cout // 256)
    kblocks = -(-cin // 32)
    taps = [(r * dil - pad, s * dil - pad) for r in range(k) for s in range(k)]

    def mask(m):
        if m >= m_tiles:
            return 0
        wt, ht = m % tw, (m // tw) % th
        mk = 0
        for i, (dh, dw) in enumerate(taps):
            lo_h, lo_w = ht * bh + dh, wt * bw + dw
            if lo_h + bh - 1 >= 0 and lo_h < h and lo_w + bw - 1 >= 0 and lo_w < w:
                mk |= 1 << i
        return mk

    pairs = -(-m_tiles // 2) * n_tiles_n
    clusters = min(74, pairs)
    loads = [0] * clusters
    for p in range(pairs):
        mp = p // n_tiles_n
        both = mask(2 * mp) | mask(2 * mp + 1)
        loads[p % clusters] += bin(both or 1).count('1') * kblocks
    lib = L.load()
    for knob in (1, 0):
        lib.b2_debug_set(17, knob)
        try:
            pl = tensor_busy.plan(n, h, w, cin, cout, k, dil)
        finally:
            lib.b2_debug_set(17, 1)
        assert (pl['bw'], pl['bh'], pl['bn']) == (bw, bh, bn)
        assert pl['tile_pairs'] == pairs and pl['cta_pairs'] == clusters
        assert pl['stages_total'] == sum(loads) and pl['stages_busiest_pair'] == max(loads)

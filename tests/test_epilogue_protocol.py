"""CPU: host-checkable parts of the round-2 kernel changes that have no arithmetic of their own.

* The operand-slot protocol of the TMA epilogue's one-operand flavours (csrc/conv_epilogue_tma.cuh, `WarpState::head` /
  `requested`): two 4 KB slots with one mbarrier each, two requests in flight, requests reaching into the next tile.  The CUDA
  code cannot run here; this is the same state machine in Python, checked for the properties the kernel relies on (every wait has
  an armed barrier, a slot is never re-armed while armed, items are consumed in order, nothing is left in flight at the end).
* The chunking of the column reductions (csrc/netops.cu `red_rows_per_chunk`) seen through b2_bn_workspace_doubles."""
import random

import pytest

from cutmix_semisup_seg_b200 import lib as L


def _drain(cnts):
    head, requested = 0, 0
    slot_item, armed, consumed = [None, None], [False, False], []
    for t, cnt in enumerate(cnts):
        nxt = cnts[t + 1] if t + 1 < len(cnts) else 0
        if cnt == 0:                                   # inactive quarter / no channel of this N tile: the warp skips the tile
            assert requested == 0
            continue

        def item(k):
            return (t, k) if k < cnt else (t + 1, k - cnt)
        while requested < 2 and requested < cnt + nxt:  # top-up at tile entry
            s = (head + requested) & 1
            assert not armed[s]
            slot_item[s], armed[s] = item(requested), True
            requested += 1
        for i in range(cnt):
            assert armed[head] and slot_item[head] == (t, i)
            armed[head] = False
            consumed.append((t, i))
            requested -= 1
            k = i + 1 + requested
            if k < cnt + nxt:
                assert not armed[head]
                slot_item[head], armed[head] = item(k), True
                requested += 1
            head ^= 1
    assert not any(armed) and requested == 0
    return consumed


def test_two_slot_operand_protocol_of_the_tma_epilogue():
    rng = random.Random(7)
    assert _drain([4, 4, 4]) == [(t, i) for t in range(3) for i in range(4)]
    for _ in range(5000):
        cnts = [rng.choice([0, 1, 2, 3, 4, 4, 4]) for _ in range(rng.randint(1, 9))]
        got = _drain(cnts)
        assert got == [(t, i) for t, c in enumerate(cnts) for i in range(c)]


@pytest.mark.parametrize('rows,c,chunks', [(262144, 256, 128),      # 2048 rows per block: 128 x 8 = 1024 blocks already
                                           (65536, 256, 128),       # 2048 would give 256 blocks = 1.7 per SM: 512 rows per block
                                           (65536, 48, 256),        # narrow tensors: 256 rows per block (the floor)
                                           (1000, 2048, 4)])        # few rows: the 256-row floor, 4 x 64 blocks
def test_column_reductions_fill_the_chip(rows, c, chunks):
    ws = int(L.call('b2_bn_workspace_doubles', rows, c))
    assert ws == chunks * c * 2 + 2 * c
    blocks = chunks * ((c + 31) // 32)
    assert blocks >= 148 * 4 or chunks == -(-rows // 256)

"""world_size=2 gloo test of the data-parallel host logic: flat gradient views + the single all-reduce(avg)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from cutmix_semisup_seg_b200.step import FlatGrads, average_gradients
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.BatchNorm2d(8), torch.nn.Conv2d(8, 4, 1))
    net[0].weight.data = net[0].weight.data.contiguous(memory_format=torch.channels_last)
    net[1].weight.requires_grad = False
    fg = FlatGrads(list(net.parameters()) + [net[0].weight])          # duplicate entry must be ignored
    ok = fg.numel == sum(p.numel() for p in net.parameters() if p.requires_grad)
    for i, p in enumerate(fg.params):
        p.grad.fill_(float(rank + 1) * (i + 1))                        # "local" gradients
        ok = ok and p.grad.stride() == p.stride()
    average_gradients(fg.flat, dist)
    for i, p in enumerate(fg.params):
        ok = ok and bool(torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1))))
    # zero_grad(set_to_none=True) drops the views; attach() must restore them onto the same storage
    for p in fg.params:
        p.grad = None
    fg.attach()
    ok = ok and all(p.grad is not None and p.grad.data_ptr() >= fg.flat.data_ptr() for p in fg.params)
    ok = ok and net[1].weight.grad is None
    # bucketed form (overlapped with the backward pass on the GPU): contiguous spans, last parameters first, shrinking towards
    # the front, covering the buffer exactly once; launching them in any grouping + finish() == the single all-reduce
    from cutmix_semisup_seg_b200.step import BucketedAllReduce
    for i, p in enumerate(fg.params):
        p.grad.fill_(float(rank + 1) * (i + 1))
    bar = BucketedAllReduce(fg, dist, None, n_buckets=3)
    spans = bar.spans
    ok = ok and spans[0][1] == fg.numel and spans[-1][0] == 0 and all(a[0] == b[1] for a, b in zip(spans, spans[1:]))
    ok = ok and sum(len(s[2]) for s in spans) == len(fg.params) and spans[0][2][0] is fg.params[-1]
    cuts = bar.ready_cuts({id(p): 10 - i for i, p in enumerate(fg.params)})
    for k in sorted(cuts, reverse=True):
        cuts[k]()
    ok = ok and bar.launched == len(spans) and bar.finish() == len(spans)
    for i, p in enumerate(fg.params):
        ok = ok and bool(torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1))))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_flat_grads_allreduce_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]

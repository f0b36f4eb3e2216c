"""Device timeline of the CUDA-graph-replayed iteration (cfg3 by default), from CUPTI through torch.profiler (no nsys in the image).

    python tools/timeline.py [arch] [out.txt]

Prints, for ONE steady-state iteration: wall time between the first kernel's start and the last kernel's end, the sum of the
kernel durations, the idle time between consecutive kernels (total, histogram) and the largest gaps with the kernels on either
side -- i.e. where the difference between `ms_per_step` and the kernel sum of the launch list goes.  Not a bench value (the
profiler adds overhead per activity record); the bench line is printed by bench.py outside any profiler."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
from cutmix_semisup_seg_b200 import synthetic  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else 'v3plus'
out_path = sys.argv[2] if len(sys.argv) > 2 else None
cfg = bench.CFG[arch]
dev = torch.device('cuda:0')
torch.cuda.set_device(dev)
trainer, mg = bench.build_trainer(cfg, dev, False, use_graph=True)
n, h, w = cfg['batch'], cfg['h'], cfg['w']
sup = synthetic.make_sup_batch(n, h, w, cfg['classes'], 100)
uns = synthetic.make_unsup_batch(n, h, w, 200, mg)
sup = tuple(t.to(dev) for t in sup)
uns = {k: v.to(dev) for k, v in uns.items()}
for _ in range(4):
    trainer.step(sup, [uns])
torch.cuda.synchronize()
STEPS = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(STEPS):
        trainer.step(sup, [uns])
    torch.cuda.synchronize()

evs = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        evs.append((e.time_range.start, e.time_range.end, e.name))
evs.sort()
lines = []
P = lines.append
P('device activities recorded: %d over %d iterations (%d per iteration)' % (len(evs), STEPS, len(evs) // STEPS))
# the middle iteration: activities are periodic, cut at equal counts
per = len(evs) // STEPS
it = evs[per:2 * per]
t0, t1 = it[0][0], max(e[1] for e in it)
busy = 0.0
cur_s, cur_e = it[0][0], it[0][1]
gaps = []
ksum = 0.0
prev = it[0]
for s, e, name in it:
    ksum += e - s
for s, e, name in it[1:]:
    if s > cur_e:
        gaps.append((s - cur_e, prev[2], name))
        busy += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
    if e >= cur_e:
        prev = (s, e, name)
busy += cur_e - cur_s
P('iteration span %.3f ms   sum of activity durations %.3f ms   device busy (union) %.3f ms   idle %.3f ms in %d gaps'
  % ((t1 - t0) / 1e3, ksum / 1e3, busy / 1e3, (t1 - t0 - busy) / 1e3, len(gaps)))
nxt = evs[2 * per][0] if len(evs) > 2 * per else t1
P('gap to the next iteration\'s first activity: %.1f us' % (nxt - t1))
edges = [1, 2, 3, 4, 6, 8, 12, 20, 50, 100, 1000, 1e9]
hist = [[0, 0.0] for _ in edges]
for g, a, b in gaps:
    for i, ed in enumerate(edges):
        if g <= ed:
            hist[i][0] += 1; hist[i][1] += g
            break
P('gap histogram (us): ' + '  '.join('<=%g: n=%d %.2f ms' % (ed, c, t / 1e3) for ed, (c, t) in zip(edges, hist) if c))


def short(nm):
    nm = nm.replace('void ', '').replace('(anonymous namespace)::', '')
    return nm.split('(')[0][:60]


P('largest gaps:')
for g, a, b in sorted(gaps, reverse=True)[:25]:
    P('  %8.1f us   after %-45s before %s' % (g, short(a), short(b)))
# idle time by the kernel that FOLLOWS the gap (launch latency of that kernel type)
by = {}
for g, a, b in gaps:
    k = short(b)
    by.setdefault(k, [0, 0.0]); by[k][0] += 1; by[k][1] += g
P('idle time by following kernel:')
for k, (c, t) in sorted(by.items(), key=lambda kv: -kv[1][1])[:30]:
    P('  %-60s n=%4d  %8.3f ms  (%.2f us each)' % (k, c, t / 1e3, t / c))
by = {}
for g, a, b in gaps:
    k = short(a)
    by.setdefault(k, [0, 0.0]); by[k][0] += 1; by[k][1] += g
P('idle time by preceding kernel:')
for k, (c, t) in sorted(by.items(), key=lambda kv: -kv[1][1])[:15]:
    P('  %-60s n=%4d  %8.3f ms  (%.2f us each)' % (k, c, t / 1e3, t / c))
# in-graph duration per kernel name (the eager per-launch profile of bench.py runs with gaps between launches, i.e. cooler)
tot = {}
for s_, e_, name in it:
    k = short(name)
    tot.setdefault(k, [0, 0.0]); tot[k][0] += 1; tot[k][1] += e_ - s_
P('in-graph time by kernel (one iteration):')
for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    P('  %-60s n=%4d  %9.3f ms  %5.1f %%  (%.1f us each)' % (k, c, t / 1e3, 100.0 * t / ksum, t / c))
txt = '\n'.join(lines)
print(txt)
if out_path:
    open(out_path, 'w').write(txt + '\n')

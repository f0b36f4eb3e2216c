"""Aggregate an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,
sm__mem_tensor_cycles_active...` launch list (tools/ncu_step.py) per kernel: launches, time, share of the
step, DRAM traffic per launch, time-weighted tensor-pipe occupancy
(sm__mem_tensor_cycles_active == tensor-pipe busy fraction of the tcgen05 kernels).  Usage: launch_list_summary.py launches.csv"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    r = csv.reader(lines)
    hdr = next(r)
    ix = {h: i for i, h in enumerate(hdr)}
    per_launch = collections.defaultdict(dict)
    names = {}
    for row in r:
        if len(row) < len(hdr):
            continue
        name = re.sub(r'\(.*', '', row[ix['Kernel Name']].replace('<unnamed>::', ''))
        name = re.sub(r'<.*', '', name).replace('void ', '').replace('(anonymous namespace)::', '')
        m, u, v = row[ix['Metric Name']], row[ix['Metric Unit']], row[ix['Metric Value']].replace(',', '')
        try:
            v = float(v)
        except ValueError:
            continue
        if m == 'gpu__time_duration.sum':
            v = v / 1e3 if u == 'ns' else (v if u == 'us' else v * 1e3)          # -> us
        if m.startswith('dram__bytes'):
            v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
        per_launch[row[ix['ID']]][m] = v
        names[row[ix['ID']]] = name
    agg = collections.defaultdict(lambda: collections.defaultdict(float))
    for i, d in per_launch.items():
        a = agg[names[i]]
        t = d.get('gpu__time_duration.sum', 0.0)
        a['n'] += 1
        a['us'] += t
        a['rd'] += d.get('dram__bytes_read.sum', 0.0)
        a['wr'] += d.get('dram__bytes_write.sum', 0.0)
        a['tensor_us'] += t * d.get('sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', d.get('sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 0.0)) / 100.0
    tot = sum(a['us'] for a in agg.values())
    print('launches {}   kernel time {:.1f} ms (per-launch times under ncu are cold-cache and serialised: read the SHARES)'
          .format(int(sum(a['n'] for a in agg.values())), tot / 1e3))
    print('{:<40s} {:>6s} {:>10s} {:>7s} {:>12s} {:>12s} {:>10s} {:>9s}'.format(
        'kernel', 'n', 'ms', 'share', 'rd MB/launch', 'wr MB/launch', 'DRAM GB/s', 'tensor %'))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
        n = a['n']
        print('{:<40s} {:6d} {:10.3f} {:6.1f}% {:12.2f} {:12.2f} {:10.0f} {:9.1f}'.format(
            k[:40], int(n), a['us'] / 1e3, 100 * a['us'] / tot, a['rd'] / n / 1e6, a['wr'] / n / 1e6,
            (a['rd'] + a['wr']) / max(a['us'], 1e-9) / 1e3, 100 * a['tensor_us'] / max(a['us'], 1e-9)))


if __name__ == '__main__':
    main(sys.argv[1])

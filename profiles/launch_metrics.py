"""Summarise an `ncu --metrics ... --csv --log-file` launch list (one row per launch x metric) per kernel and per step.
usage: python profiles/launch_metrics.py launches.csv STEPS [key] > summary.txt
With `key` (e.g. 300_bf16_b32) also updates profiles/conv_dram_traffic.json with the conv family's DRAM bytes per step
(the `traffic` field of bench.py's roofline object)."""
import collections
import csv
import json
import os
import sys


def main(path, steps, key=None):
    rows = list(csv.reader(l for l in open(path, errors='replace') if l.startswith('"')))
    hdr = rows[0]
    iid, ikn, imn, imu, imv = (hdr.index(k) for k in ('ID', 'Kernel Name', 'Metric Name', 'Metric Unit', 'Metric Value'))
    launches = collections.OrderedDict()
    for r in rows[1:]:
        d = launches.setdefault(r[iid], {'kernel': r[ikn]})
        try:
            v = float(r[imv].replace(',', ''))
        except ValueError:
            continue
        u = r[imu]
        if u in ('ns', 'nsecond'):
            v /= 1e3
        elif u in ('ms', 'msecond'):
            v *= 1e3
        elif u in ('Kbyte', 'KB'):
            v *= 1e3
        elif u in ('Mbyte', 'MB'):
            v *= 1e6
        elif u in ('Gbyte', 'GB'):
            v *= 1e9
        d[r[imn]] = v
    agg = collections.OrderedDict()
    for d in launches.values():
        k = d['kernel'].split('(')[0].replace('void ', '').replace('ctx::', '')[:60]
        a = agg.setdefault(k, collections.Counter())
        a['n'] += 1
        a['us'] += d.get('gpu__time_duration.sum', 0.0)
        a['rd'] += d.get('dram__bytes_read.sum', 0.0)
        a['wr'] += d.get('dram__bytes_write.sum', 0.0)
        a['tensor_w'] += d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 0.0) * d.get('gpu__time_duration.sum', 0.0)
        a['l2_w'] += d.get('lts__throughput.avg.pct_of_peak_sustained_elapsed', 0.0) * d.get('gpu__time_duration.sum', 0.0)
    tot = sum(a['us'] for a in agg.values())
    print('# %s: %d launches over %d steps; per-step figures (cold-cache, serialised by ncu: shares, not absolute times)' % (os.path.basename(path), len(launches), steps))
    print('%-62s %8s %10s %7s %10s %10s %8s %6s' % ('kernel', 'launches', 'us/step', 'share', 'rd MB/step', 'wr MB/step', 'tensor%', 'L2%'))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
        print('%-62s %8.1f %10.1f %6.1f%% %10.1f %10.1f %8.1f %6.1f' % (k, a['n'] / steps, a['us'] / steps, 100 * a['us'] / tot, a['rd'] / steps / 1e6,
                                                                       a['wr'] / steps / 1e6, a['tensor_w'] / max(a['us'], 1e-9), a['l2_w'] / max(a['us'], 1e-9)))
    print('%-62s %8.1f %10.1f' % ('TOTAL', len(launches) / steps, tot / steps))
    if key:
        conv = [a for k, a in agg.items() if k.startswith('conv_')]
        out = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'conv_dram_traffic.json')
        tj = json.load(open(out)) if os.path.exists(out) else {}
        tj[key] = {'dram_bytes_per_step': (sum(a['rd'] for a in conv) + sum(a['wr'] for a in conv)) / steps,
                   'conv_launches_per_step': sum(a['n'] for a in conv) / steps,
                   'source': 'profiles/%s (ncu dram__bytes_read.sum + dram__bytes_write.sum summed over the conv launches of one step)' % os.path.basename(path)}
        json.dump(tj, open(out, 'w'), indent=1)


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]), sys.argv[3] if len(sys.argv) > 3 else None)

"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of step)."""
import collections
import csv
import sys


def main(path, skip_launches=0):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[1 + skip_launches:]:
        try:
            v = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        us = v / 1000.0 if r[ui] in ('ns', 'nsecond') else v
        name = r[ki].split('(')[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print('%-70s %8s %12s %7s' % ('kernel', 'launches', 'total us', 'share'))
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-70s %8d %12.1f %6.1f%%' % (name[:70], n, us, 100 * us / tot))
    print('%-70s %8d %12.1f' % ('TOTAL', sum(a[0] for a in agg.values()), tot))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)

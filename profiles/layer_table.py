"""Print the per-op timing table bench.py --layers wrote (CUDA events around every op, min of 3)."""
import collections
import json
import sys


def main(path, top=30):
    d = json.load(open(path))
    ops = d['ops']
    tot = sum(o['ms'] for o in ops)
    print('precision %s batch %d: %d ops, %.3f ms summed' % (d['precision'], d['batch'], len(ops), tot))
    for o in sorted(ops, key=lambda o: -o['ms'])[:top]:
        tf = o['gflop'] / o['ms'] if o['ms'] > 0 else 0
        print('%-26s %-10s %8.3f ms %5.1f%% %8.1f GF %7.1f TF/s %s' % (o['op'], o['kind'], o['ms'], 100 * o['ms'] / tot, o['gflop'], tf, o['shape']))
    k = collections.defaultdict(float)
    for o in ops:
        k[o['kind']] += o['ms']
    print({a: round(b, 3) for a, b in k.items()})


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)

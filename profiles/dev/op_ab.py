"""Development aid: selected ops of the real 300x300 bf16 engine, isolated (cold: L2 flushed before each launch; warm: 20 launches back to
back) and the graph step.  Same-box A/B: run it several times with an environment switch set both ways.
usage: op_ab.py <name substring> [<name substring> ...]"""
import os
import sys

import torch

sys.path.insert(0, '.')
import bench  # noqa


def main(names):
    dev = torch.device('cuda:0')
    torch.cuda.set_device(dev)
    w = bench.Workload(300, 'bf16', 32, dev, 'hard', 0)
    eng = w.eng
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(5):
        eng.load_input(w.x_dev); eng.launch()
    torch.cuda.synchronize()
    idx = [i for i, l in enumerate(eng.layers) if any(n in l[0] for n in names)]
    out = []
    for i in idx:
        cold, warm = [], []
        for _ in range(7):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); eng.run_range(i, i + 1, 1); b.record(); b.synchronize()
            cold.append(a.elapsed_time(b))
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); eng.run_range(i, i + 1, 20); b.record(); b.synchronize()
            warm.append(a.elapsed_time(b) / 20)
        out.append('%s cold %.1f warm %.1f us' % (eng.layers[i][0][:24], 1e3 * min(cold), 1e3 * min(warm)))
    ts = []
    for _ in range(40):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.load_input(w.x_dev); eng.launch(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print('step median %.4f min %.4f ms | %s' % (ts[len(ts) // 2], ts[0], ' | '.join(out)))


if __name__ == '__main__':
    main(sys.argv[1:] or ['ConvLinear'])

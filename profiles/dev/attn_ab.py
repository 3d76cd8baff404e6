"""Development aid: the Context-Transformer kernel inside the real 300x300 engine, isolated (20 launches, one event pair) and the graph
step, for the current setting of CTX_ATTN_QT1 (read once per process).  Same-box A/B: run twice with the switch set both ways."""
import os
import sys

import torch

sys.path.insert(0, '.')
import bench  # noqa


def main():
    dev = torch.device('cuda:0')
    torch.cuda.set_device(dev)
    w = bench.Workload(300, 'bf16', 32, dev, 'hard', 0)
    eng = w.eng
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(5):
        eng.load_input(w.x_dev); eng.launch()
    torch.cuda.synchronize()
    idx = [i for i, l in enumerate(eng.layers) if l[1] == 'attention']
    res = {}
    for i in idx:
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); eng.run_range(i, i + 1, 20); b.record(); b.synchronize()
            ts.append(a.elapsed_time(b) / 20)
        res[eng.layers[i][0]] = min(ts)
    ts = []
    for _ in range(30):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.load_input(w.x_dev); eng.launch(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print('CTX_ATTN_QT1=%s attention %s  step median %.4f min %.4f ms' % (os.environ.get('CTX_ATTN_QT1', '(default)'), res, ts[len(ts) // 2], ts[0]))


if __name__ == '__main__':
    main()

"""Development aid: per-op warm times (20 launches back to back) of an engine, sorted (default: 512x512 fp16 B=16, BASELINE config 3; LAYERS_SIZE / LAYERS_PRECISION / LAYERS_BATCH select another)."""
import sys

import torch

sys.path.insert(0, '.')
import bench  # noqa


def main():
    dev = torch.device('cuda:0')
    torch.cuda.set_device(dev)
    import os
    size, prec, batch = (int(os.environ.get("LAYERS_SIZE", "512")), os.environ.get("LAYERS_PRECISION", "fp16"), int(os.environ.get("LAYERS_BATCH", "16")))
    w = bench.Workload(size, prec, batch, dev, "linear" if size == 512 else "hard", 0)
    eng = w.eng
    for _ in range(3):
        eng.load_input(w.x_dev); eng.launch()
    torch.cuda.synchronize()
    rows = []
    for i, l in enumerate(eng.layers):
        ts = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); eng.run_range(i, i + 1, 20); b.record(); b.synchronize()
            ts.append(a.elapsed_time(b) / 20)
        rows.append((min(ts), i, l[0], l[1], l[2] / 1e9, l[3], eng.conv_config(i) if l[1].startswith('conv') else None))
    print('sum of warm op times %.3f ms over %d ops' % (sum(r[0] for r in rows), len(rows)))
    for r in sorted(rows, reverse=True)[:45]:
        print('%2d %-34s %-8s %.4f ms %6.1f GF %5.0f TF/s %s %s' % (r[1], r[2][:34], r[3], r[0], r[4], r[4] / r[0] if r[0] else 0, list(r[5])[:9], r[6]))


if __name__ == '__main__':
    main()

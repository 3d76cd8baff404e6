"""Development aid: candidate / kept list lengths per (image, class) of the 512x512 soft-NMS workload (BASELINE config 3) and the
time of the post chain alone."""
import sys

import torch

sys.path.insert(0, '.')
import bench  # noqa


def main():
    dev = torch.device('cuda:0')
    torch.cuda.set_device(dev)
    w = bench.Workload(512, 'fp16', 16, dev, 'linear', 0)
    with torch.no_grad():
        pred = w.net(w.x_dev)
    pred = [t.float().clone() for t in pred]
    post = w.post
    for _ in range(3):
        rec, cnt, _ = post(pred, w.priors, w.scale)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); post(pred, w.priors, w.scale); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    B, P, C = 16, w.priors.size(0), 20
    off = (16 * B * P + 255) // 256 * 256
    counts = post._ws[off:off + 4 * B * C * 2].view(torch.int32).cpu().view(2, B, C)
    cand, kept = counts[0].flatten().sort().values, counts[1].flatten().sort().values
    q = lambda t: [int(t[int(f * (len(t) - 1))]) for f in (0, 0.25, 0.5, 0.75, 0.9, 1.0)]
    print('post chain %.1f us/img (min of 10), P %d' % (1e3 * min(ts) / B, P))
    print('candidates per (image, class): min/25/50/75/90/max', q(cand), 'sum', int(cand.sum()))
    print('kept       per (image, class): min/25/50/75/90/max', q(kept), 'sum', int(kept.sum()))
    print('detections per image', cnt.cpu().tolist())


if __name__ == '__main__':
    main()

"""Per-kernel durations of the post-processing chain (run under ncu --metrics gpu__time_duration.sum)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
size, prec, B, nms = int(sys.argv[1]), sys.argv[2], int(sys.argv[3]), sys.argv[4]
dev = torch.device('cuda', 0)
w = bench.Workload(size, prec, B, dev, nms, seed=0)
pred = w.net(w.x_dev)
for _ in range(2):
    rec, cnt, _ = w.post.forward(pred, w.priors, w.scale)
torch.cuda.synchronize()
torch.cuda.profiler.start()
rec, cnt, _ = w.post.forward(pred, w.priors, w.scale)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('detections', int(cnt.sum()))
import context_transformer_b200 as ctx
bx, sc = ctx.Detect(21, 0, w.cfg).forward(pred, w.priors)
n = (sc[:, :, 1:] > 0.01).sum(1)          # [B, 20] candidates per (image, class)
print('candidates per (image, class): min %d mean %.0f max %d; kept per image %s' % (int(n.min()), float(n.float().mean()), int(n.max()), cnt.tolist()[:4]))

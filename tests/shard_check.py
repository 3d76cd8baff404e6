"""torchrun helper (not collected): 2 ranks, batch 4 sharded 2+2 + one NCCL all-gather must equal the
single-GPU result bit for bit."""
import os
import sys
import types

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import context_transformer_b200 as ctx
from context_transformer_b200 import shard
from oracle import synth


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl')
    dev = 'cuda:%d' % torch.cuda.current_device()
    net = ctx.build_net(types.SimpleNamespace(method='ft', phase=2, setting='transfer'), 300, 20)
    net.load_state_dict(synth.seeded_state(net.state_dict(), seed=0))
    net.eval()
    net.device = dev
    net.cuda()
    priors = ctx.PriorBox(ctx.VOC_300).forward().cuda()
    post = ctx.DetectPost(21, 0, ctx.VOC_300, score_thresh=0.3)
    scale = np.array([500, 375, 500, 375], np.float32)
    x = synth.seeded_input(4, 300, seed=2)
    lo, hi = shard.shard_bounds(4, world, rank)
    det = shard.ShardedDetector(net, post, priors)
    rec, cnt = det(x[lo:hi], scale)
    rec, cnt = rec.clone(), cnt.clone()
    net.invalidate_engine()
    rec1, cnt1 = det.local(x, scale)
    ok = torch.equal(rec, rec1) and torch.equal(cnt, cnt1) and int(cnt.sum()) > 0
    # fine-tune loss under data parallelism (BASELINE config 5): world-scaled shard losses with the all-reduced N average
    # to the single-process loss of the whole batch (multibox_loss_combined.py:119-122)
    g = torch.Generator().manual_seed(11)
    P = priors.size(0)
    loc_p, conf_p, obj_p = (torch.randn(4, P, 4, generator=g).cuda(), torch.randn(4, P, 20, generator=g).cuda(),
                            torch.randn(4, P, 2, generator=g).cuda())
    targets = []
    for b in range(4):
        n = 1 + b
        xy = torch.rand(n, 2, generator=g) * 0.5
        wh = 0.1 + torch.rand(n, 2, generator=g) * 0.4
        lab = torch.randint(1, 21, (n, 1), generator=g).float()
        targets.append(torch.cat([xy, xy + wh, lab, torch.ones(n, 1)], 1))
    crit = ctx.MultiBoxLoss_combined(21, 0.5, True, 0, True, 3, 0.5, False)
    mine = crit((loc_p[lo:hi], conf_p[lo:hi], obj_p[lo:hi]), priors, targets[lo:hi])
    vec = torch.stack([mine['loss_box_reg'], mine['loss_cls'], mine['loss_obj']]).double()
    dist.all_reduce(vec)
    vec /= world
    crit.process_group = None
    dist_initialised = dist.is_initialized
    dist.is_initialized = lambda: False                       # single-process semantics for the reference value
    try:
        full = crit((loc_p, conf_p, obj_p), priors, targets)
    finally:
        dist.is_initialized = dist_initialised
    want = torch.stack([full['loss_box_reg'], full['loss_cls'], full['loss_obj']]).double()
    loss_ok = bool(torch.allclose(vec, want, rtol=1e-5, atol=1e-6))
    ok = ok and loss_ok
    flag = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print('SHARD_OK' if int(flag) else 'SHARD_MISMATCH', cnt.tolist(), cnt1.tolist(), 'loss', vec.tolist(), want.tolist())
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == '__main__':
    main()

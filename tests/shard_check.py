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
    flag = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print('SHARD_OK' if int(flag) else 'SHARD_MISMATCH', cnt.tolist(), cnt1.tolist())
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == '__main__':
    main()

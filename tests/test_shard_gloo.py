"""CPU, world_size 2 over gloo: the N>1 exchange step (one all-gather of packed records)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from context_transformer_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        B_global, K = 6, 9
        g = torch.Generator().manual_seed(1234)
        rec_all = torch.randn(B_global, K, 6, generator=g)
        cnt_all = torch.randint(0, K + 5, (B_global,), generator=g, dtype=torch.int32)
        lo, hi = shard.shard_bounds(B_global, world, rank)
        rec, cnt = shard.gather_records(rec_all[lo:hi].clone(), cnt_all[lo:hi].clone())
        ok = torch.equal(rec, rec_all) and torch.equal(cnt, cnt_all)
        # loss normaliser of the data-parallel fine-tune loop: N is the positive count of the GLOBAL batch, and the
        # world-scaled shard losses average (DDP) to global_sum / N_global
        num_pos = torch.tensor([7, 0, 3, 11, 5, 2])
        sums = torch.tensor([1.5, 0.0, 0.25, 4.0, 2.0, 0.5], dtype=torch.float64)
        n, w = shard.global_positive_count(num_pos[lo:hi].sum())
        mine = sums[lo:hi].sum() * w / n
        tot = mine.clone()
        dist.all_reduce(tot)
        ok = ok and int(n) == 28 and w == world and abs(float(tot) / world - float(sums.sum()) / 28.0) < 1e-12
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_gather_records_world2():
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}

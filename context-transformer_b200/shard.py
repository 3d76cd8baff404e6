"""Multi-GPU batched inference: shard the batch, gather the detections once.

The path shards naturally — every image is independent through the conv stack, the per-image
attention softmax, decode and NMS — so each rank (one process per GPU,
``torch.distributed``) runs the compiled network and ``DetectPost`` on its contiguous slice of the
global batch with no data-path collective, and the results are exchanged with ONE
``all_gather_into_tensor`` of fixed-shape records (NCCL over NVLink/NVSwitch on GPUs; the same code
runs over ``gloo`` in the CPU tests).  The reference has no multi-GPU inference at all
(test.py is batch-1, single device); its only parallelism is ``nn.DataParallel`` in training.

Wire format per image: ``[max_out + 1, 6]`` float32 — rows 0..max_out-1 are the detection records
(x1,y1,x2,y2,score,class), the last row is ``[count, 0, 0, 0, 0, 0]`` — 4.8 KB at max_out = 200+.
"""
import torch
import torch.distributed as dist


def shard_bounds(global_batch, world_size, rank):
    """Contiguous, equal shards (all_gather_into_tensor needs equal sizes)."""
    if global_batch % world_size:
        raise ValueError('global batch %d is not divisible by world size %d' % (global_batch, world_size))
    per = global_batch // world_size
    return rank * per, (rank + 1) * per


def pack_records(records, counts):
    """records[B,K,6] f32 + counts[B] int -> packed[B,K+1,6] f32."""
    B, K, _ = records.shape
    packed = torch.zeros(B, K + 1, 6, dtype=torch.float32, device=records.device)
    packed[:, :K] = records
    packed[:, K, 0] = counts.to(torch.float32)
    return packed


def unpack_records(packed):
    K = packed.size(1) - 1
    return packed[:, :K], packed[:, K, 0].round().to(torch.int32)


def gather_records(records, counts, group=None):
    """One all-gather of the packed records; returns (records[B_global,K,6], counts[B_global])
    ordered by rank, i.e. in global batch order for contiguous shards."""
    packed = pack_records(records, counts)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return unpack_records(packed)
    world = dist.get_world_size(group)
    out = torch.empty((world * packed.size(0),) + tuple(packed.shape[1:]), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out, packed.contiguous(), group=group)
    return unpack_records(out)


def global_positive_count(num_pos_local, group=None):
    """Data-parallel fine-tuning (BASELINE config 5): the reference normalises every loss term by the positive count of
    the WHOLE batch (``N = num_pos.sum()`` after ``nn.DataParallel`` has gathered the replicas' outputs on GPU 0,
    multibox_loss_combined.py:119-122).  With one process per GPU each rank only sees its shard, so N needs one scalar
    all-reduce.  Returns ``(N_global, world)``: a rank's loss is ``local_sum * world / N_global`` — gradient averaging
    over ranks (DDP) then yields exactly ``global_sum / N_global``."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return num_pos_local, 1
    n = num_pos_local.detach().clone()
    dist.all_reduce(n, op=dist.ReduceOp.SUM, group=group)
    return n, dist.get_world_size(group)


class ShardedDetector(object):
    """net: RFBNet in eval mode on this rank's GPU; post: DetectPost; priors: [P,4]."""

    def __init__(self, net, post, priors, group=None):
        self.net, self.post, self.priors, self.group = net, post, priors, group

    def local(self, x_local, scale):
        pred = self.net(x_local)
        records, counts, _ = self.post.forward(pred, self.priors, scale)
        return records, counts

    def forward(self, x_local, scale):
        records, counts = self.local(x_local, scale)
        return gather_records(records, counts, self.group)

    __call__ = forward

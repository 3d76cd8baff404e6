"""``MultiBoxLoss_combined`` — training loss of the few-shot fine-tune loop.

Mirror of reference ``layers/modules/multibox_loss_combined.py:7-124`` (same constructor and
``forward(predictions, priors, targets) -> {'loss_box_reg', 'loss_cls', 'loss_obj'}``), evaluated by four kernel
launches for the whole batch instead of ~40 framework ops with boolean-mask gathers and an autograd tape:

* ``ctx_match_encode``           target assignment (``match`` + ``encode``, the Python loop of :70-74)
* ``ctx_loss_mining``            the no-grad objectness cross-entropy that drives mining (:88-90) + the positive count (:77)
* ``ctx_hard_negative_rank``     the two full sorts of :91-93
* ``ctx_loss_forward_backward``  smooth-L1, objectness CE and the logit-combined class CE (:81-117) AND their gradients
                                 w.r.t. loc / conf / obj in one pass (``_FusedLoss`` hands them to autograd)

Under ``torch.distributed`` (one process per GPU) the normaliser N is all-reduced over the replicas
(see ``shard.global_positive_count``).
"""
import torch
import torch.nn as nn

from . import _lib
from .box_utils import hard_negative_rank, match_batch
from .shard import global_positive_count


class _FusedLoss(torch.autograd.Function):
    """(loc, conf, obj) -> the three un-normalised loss sums; backward returns the gradients the kernel wrote alongside."""

    @staticmethod
    def forward(ctx, loc, conf, obj, loc_t, conf_t, obj_u8, rank, num_neg):
        B, P, C = conf.shape
        dev = loc.device
        locc, confc, objc = (t.detach().float().contiguous() for t in (loc, conf, obj))
        sums = torch.empty(3, dtype=torch.float64, device=dev)
        dloc = torch.empty(B, P, 4, device=dev)
        dconf = torch.empty(B, P, C, device=dev)
        dobj_c = torch.empty(B, P, 2, device=dev)
        dobj_o = torch.empty(B, P, 2, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().ctx_loss_forward_backward(
                locc.data_ptr(), confc.data_ptr(), objc.data_ptr(), loc_t.data_ptr(), conf_t.data_ptr(), obj_u8.data_ptr(),
                rank.data_ptr(), num_neg.data_ptr(), B, P, C, sums.data_ptr(), dloc.data_ptr(), dconf.data_ptr(),
                dobj_c.data_ptr(), dobj_o.data_ptr(), _lib.current_stream_ptr()), 'ctx_loss_forward_backward')
        ctx.save_for_backward(dloc, dconf, dobj_c, dobj_o)
        ctx.in_dtypes = (loc.dtype, conf.dtype, obj.dtype)
        out = sums.float()
        return out[0], out[1], out[2]

    @staticmethod
    def backward(ctx, g_l, g_c, g_o):
        dloc, dconf, dobj_c, dobj_o = ctx.saved_tensors
        tl, tc, to = ctx.in_dtypes
        return ((dloc * g_l).to(tl), (dconf * g_c).to(tc), (dobj_c * g_c + dobj_o * g_o).to(to), None, None, None, None, None)


class MultiBoxLoss_combined(nn.Module):
    def __init__(self, num_classes, overlap_thresh, prior_for_matching, bkg_label, neg_mining, neg_pos,
                 neg_overlap, encode_target):
        super(MultiBoxLoss_combined, self).__init__()
        self.num_classes = num_classes
        self.threshold = overlap_thresh
        self.background_label = bkg_label
        self.encode_target = encode_target
        self.use_prior_for_matching = prior_for_matching
        self.do_neg_mining = neg_mining
        self.negpos_ratio = neg_pos
        self.neg_overlap = neg_overlap
        self.variance = [0.1, 0.2]
        self.process_group = None          # torch.distributed group of the data-parallel replicas (None = default group)

    def forward(self, predictions, priors, targets):
        loc_data, conf_data, obj_data = predictions
        dev = _lib.require_cuda(loc_data, 'loc predictions').device
        B, P = loc_data.size(0), priors.size(0)
        if conf_data.size(-1) != self.num_classes - 1:
            raise ValueError('MultiBoxLoss_combined: conf has %d channels, expected num_classes - 1 = %d' % (conf_data.size(-1), self.num_classes - 1))
        loc_t, conf_t, obj_t, _ = match_batch(self.threshold, targets, priors.detach().to(dev), self.variance, obj_as_u8=True)
        L = _lib.lib()
        mining = torch.empty(B, P, device=dev)
        num_pos_w = torch.empty(B, dtype=torch.float64, device=dev)
        objc = obj_data.detach().float().contiguous()
        with torch.cuda.device(dev):
            _lib.check(L.ctx_loss_mining(objc.data_ptr(), conf_t.data_ptr(), obj_t.data_ptr(), B, P, mining.data_ptr(), num_pos_w.data_ptr(),
                                         _lib.current_stream_ptr()), 'ctx_loss_mining')
        num_pos = num_pos_w.long()                                                 # (weights * pos).sum(1).long(), :77
        num_neg = torch.clamp(self.negpos_ratio * num_pos, max=P - 1).contiguous()                    # :94
        rank = hard_negative_rank(mining)
        sum_l, sum_c, sum_o = _FusedLoss.apply(loc_data, conf_data.reshape(B, P, -1), obj_data, loc_t, conf_t, obj_t, rank, num_neg)
        # single process: N = positives of the batch, as upstream (:119).  One process per GPU (torch.distributed initialised):
        # N is the positive count of the global batch (one scalar all-reduce) and the shard's sums are scaled by the world
        # size, so that DDP's gradient averaging reproduces the reference's DataParallel loss exactly.
        N, world = global_positive_count(num_pos.sum(), self.process_group)
        k = float(world)
        return {'loss_box_reg': sum_l * k / N, 'loss_cls': sum_c * k / N, 'loss_obj': sum_o * k / N}

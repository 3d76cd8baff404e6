"""``MultiBoxLoss_combined`` — training loss of the few-shot fine-tune loop.

Mirror of reference ``layers/modules/multibox_loss_combined.py:7-124`` (same constructor and
``forward(predictions, priors, targets) -> {'loss_box_reg', 'loss_cls', 'loss_obj'}``).  The two
non-differentiable, data-dependent stages run as CUDA kernels for the whole batch:

* target assignment (``match`` + ``encode``, the Python loop of :70-74)  -> ``ctx_match_encode``
* hard-negative ranking (the two full sorts of :91-93)                    -> ``ctx_hard_negative_rank``

The differentiable reductions (smooth-L1, the two cross-entropies on the mined set, the
logit-combine of :106-117) stay as autograd tensor expressions on the same device so that
``loss.backward()`` reaches the network exactly as upstream.  Under ``torch.distributed`` (one process
per GPU) the normaliser N is all-reduced over the replicas (see ``shard.global_positive_count``).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .box_utils import hard_negative_rank, match_batch
from .shard import global_positive_count


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
        device = loc_data.device
        num = loc_data.size(0)
        num_priors = priors.size(0)

        loc_t, conf_t, obj_t, _ = match_batch(self.threshold, targets, priors.detach().to(device), self.variance)

        pos = conf_t[:, :, 0] > 0
        num_pos = (conf_t[:, :, 1] * pos.float()).sum(1, keepdim=True).long()

        loss_l = F.smooth_l1_loss(loc_data[pos], loc_t[pos], reduction='none')
        weight_pos = conf_t[pos][:, 1]
        loss_l = torch.sum(torch.sum(loss_l, dim=1) * weight_pos)

        with torch.no_grad():
            mining = F.cross_entropy(obj_data.reshape(-1, 2), obj_t.long().view(-1), reduction='none')
            mining[obj_t.view(-1)] = 0
            idx_rank = hard_negative_rank(mining.view(num, -1))
            num_neg = torch.clamp(self.negpos_ratio * num_pos, max=num_priors - 1)
            neg = idx_rank < num_neg.expand_as(idx_rank)

        mask = pos | neg
        weight = conf_t[mask][:, 1]
        loss_obj = torch.sum(F.cross_entropy(obj_data[mask], obj_t[mask].long(), reduction='none') * weight)

        batch_conf = conf_data.reshape(-1, self.num_classes - 1)
        batch_obj = obj_data.reshape(-1, 2)
        logit_0 = batch_obj[:, 0].unsqueeze(1) + torch.log(torch.exp(batch_conf).sum(dim=1, keepdim=True))
        logit_k = batch_obj[:, 1].unsqueeze(1).expand_as(batch_conf) + batch_conf
        logit = torch.cat((logit_0, logit_k), 1).view(num, -1, self.num_classes)
        loss_c = torch.sum(F.cross_entropy(logit[mask], conf_t[mask][:, 0].long(), reduction='none') * weight)

        # single process: N = positives of the batch, as upstream.  One process per GPU (torch.distributed initialised):
        # N is the positive count of the global batch (one scalar all-reduce) and the shard's sums are scaled by the world
        # size, so that DDP's gradient averaging reproduces the reference's DataParallel loss exactly.
        N, world = global_positive_count(num_pos.sum(), self.process_group)
        k = float(world)
        return {'loss_box_reg': loss_l * k / N, 'loss_cls': loss_c * k / N, 'loss_obj': loss_obj * k / N}

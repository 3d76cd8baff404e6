"""``init_reweight`` — OBJ(Target) prototype initialisation of the few-shot fine-tune loop.

Mirror of reference ``train.py:252-286``: for up to ``init_iter`` batches, take the raw conf features
``model(data, init=True)`` ([B, P, C_src]), match the priors to the ground truth, collect per foreground
class the features of its positive priors, L2-normalise each row, average per class, normalise the mean
and write the result into ``model.OBJ_Target.weight``.

Upstream does the collection with 20 boolean-mask gathers and growing ``torch.cat`` lists per batch
(dynamic shapes, one sync each).  Here target assignment is the batched ``ctx_match_encode`` kernel and
the collection is ONE pass per batch (``ctx_prototype_accumulate``: normalised rows added into a
per-class fp64 sum) plus a final ``ctx_prototype_finalize`` — no dynamic shapes, no host sync until the
weights are read.  The forward itself (``init=True`` in ``train()`` mode uses batch-statistics BatchNorm)
stays the module's autograd path, as for the rest of training.
"""
import torch

from . import _lib
from .box_utils import match_batch


class PrototypeAccumulator(object):
    """Running per-class sums of L2-normalised feature rows (train.py:268-279)."""

    def __init__(self, num_fg, dim, device):
        self.num_fg, self.dim = int(num_fg), int(dim)
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise _lib.CtxError('PrototypeAccumulator needs a CUDA device (no CPU fallback), got %s' % device)
        self.sums = torch.zeros(self.num_fg, self.dim, dtype=torch.float64, device=self.device)
        self.counts = torch.zeros(self.num_fg, dtype=torch.int32, device=self.device)

    def add(self, feat, conf_t):
        """feat [B,P,dim] fp32 (raw conf features), conf_t [B,P,2] fp32 (label, weight) from ``match``."""
        feat = _lib.require_cuda(feat, 'feat').detach().float().contiguous()
        conf_t = _lib.require_cuda(conf_t, 'conf_t').float().contiguous()
        B, P, D = feat.shape
        if D != self.dim or tuple(conf_t.shape) != (B, P, 2):
            raise ValueError('feat %s / conf_t %s do not match dim %d' % (tuple(feat.shape), tuple(conf_t.shape), self.dim))
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ctx_prototype_accumulate(feat.data_ptr(), conf_t.data_ptr(), B, P, D, self.num_fg,
                                                          self.sums.data_ptr(), self.counts.data_ptr(),
                                                          _lib.current_stream_ptr()), 'ctx_prototype_accumulate')

    def prototypes(self, first_class=0):
        """[num_fg - first_class, dim] fp32: normalise(mean of normalised rows) per class (NaN for a class without samples)."""
        out = torch.empty(self.num_fg - first_class, self.dim, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ctx_prototype_finalize(self.sums.data_ptr(), self.counts.data_ptr(), self.num_fg, self.dim,
                                                        int(first_class), out.data_ptr(), _lib.current_stream_ptr()),
                       'ctx_prototype_finalize')
        return out


def init_reweight(args, model, data_loader, priors, num_classes=21, overlap_threshold=0.5, variances=(0.1, 0.2)):
    """Reference call ``init_reweight(args, model, data_loader)`` (train.py:188); ``priors``, ``num_classes`` and
    ``overlap_threshold`` are module globals upstream (train.py:137-141, :70, :66).  ``args`` needs ``init_iter`` and
    ``setting``.  ``data_loader`` yields ``(data[B,3,S,S], targets: list of [n_i, 6])`` like ``detection_collate``."""
    net = model.module if hasattr(model, 'module') else model
    dev = next(net.parameters()).device
    acc = None
    for (data, targets), _ in zip(data_loader, range(args.init_iter)):
        with torch.no_grad():
            conf_data = model(data.to(dev), init=True)                      # [B, P, C_src]
        _, conf_t, _, _ = match_batch(overlap_threshold, targets, priors.to(dev), variances)
        if acc is None:
            acc = PrototypeAccumulator(num_classes - 1, conf_data.size(-1), dev)
        acc.add(conf_data, conf_t)
    if acc is None:
        raise ValueError('init_reweight: the data loader produced no batch')
    first = 15 if getattr(args, 'setting', 'transfer') == 'incre' else 0   # train.py:281-282
    net.OBJ_Target.weight.data = acc.prototypes(first)
    return net.OBJ_Target.weight.data

"""Box algebra of the detection path — mirror of the live part of reference ``utils/box_utils.py``.

``match`` (box_utils.py:83-132) keeps the reference's in-place signature but runs on the GPU
(``ctx_match_encode``); ``match_batch`` is the batched form the loss uses (one launch for the whole
batch instead of the Python loop of multibox_loss_combined.py:70-74).  ``decode`` (:184-202) goes
through the ``Detect`` kernel.  ``point_form`` / ``jaccard`` / ``encode`` call the stand-alone device ops that share
the match kernel's device functions (``ctx_point_form`` / ``ctx_jaccard`` / ``ctx_encode``).
"""
import torch

from . import _lib


def _boxes(t, name):
    t = _lib.require_cuda(t, name).detach().float().contiguous()
    if t.dim() != 2 or t.size(1) != 4:
        raise ValueError('%s must be [n,4]' % name)
    return t


def point_form(boxes):
    """(cx, cy, w, h) -> corners (box_utils.py:5-14), ``ctx_point_form``."""
    boxes = _boxes(boxes, 'boxes')
    out = torch.empty_like(boxes)
    with torch.cuda.device(boxes.device):
        _lib.check(_lib.lib().ctx_point_form(boxes.data_ptr(), boxes.size(0), out.data_ptr(), _lib.current_stream_ptr()), 'ctx_point_form')
    return out


def jaccard(box_a, box_b):
    """IoU matrix [A,B] of corner-form boxes (box_utils.py:50-68), ``ctx_jaccard`` — the device function the match kernel uses."""
    box_a, box_b = _boxes(box_a, 'box_a'), _boxes(box_b, 'box_b')
    out = torch.empty(box_a.size(0), box_b.size(0), device=box_a.device)
    with torch.cuda.device(box_a.device):
        _lib.check(_lib.lib().ctx_jaccard(box_a.data_ptr(), box_a.size(0), box_b.data_ptr(), box_b.size(0), out.data_ptr(),
                                          _lib.current_stream_ptr()), 'ctx_jaccard')
    return out


def encode(matched, priors, variances):
    """Regression targets of matched corner boxes w.r.t. centre-form priors (box_utils.py:135-156), ``ctx_encode``."""
    matched, priors = _boxes(matched, 'matched'), _boxes(priors, 'priors')
    if matched.size(0) != priors.size(0):
        raise ValueError('encode: %d matched boxes for %d priors' % (matched.size(0), priors.size(0)))
    out = torch.empty_like(matched)
    with torch.cuda.device(matched.device):
        _lib.check(_lib.lib().ctx_encode(matched.data_ptr(), priors.data_ptr(), matched.size(0), float(variances[0]), float(variances[1]),
                                         out.data_ptr(), _lib.current_stream_ptr()), 'ctx_encode')
    return out


def decode(loc, priors, variances):
    """[P,4] offsets + [P,4] priors -> [P,4] corner boxes (box_utils.py:184-202), on the GPU."""
    from .detection import Detect
    P = loc.size(0)
    dummy_conf = torch.zeros(1, P, 1, device=loc.device)
    dummy_obj = torch.zeros(1, P, 2, device=loc.device)
    boxes, _ = Detect(2, 0, {'variance': variances}).forward((loc.unsqueeze(0), dummy_conf, dummy_obj), priors)
    return boxes[0]


def match_batch(threshold, targets, priors, variances, want_overlap=False, obj_as_u8=False):
    """Batched ``match``: targets is a list of [n_i,6] tensors (x1,y1,x2,y2,label,weight).
    Returns loc_t[B,P,4] f32, conf_t[B,P,2] f32, obj_t[B,P] bool, best_truth_idx[B,P] int32
    (and best_truth_overlap[B,P] when asked)."""
    priors = _lib.require_cuda(priors, 'priors').float().contiguous()
    dev = priors.device
    B, P = len(targets), priors.size(0)
    max_obj = max([int(t.size(0)) for t in targets] + [1])
    nobj = torch.tensor([int(t.size(0)) for t in targets], dtype=torch.int32)
    if all(t.is_cuda for t in targets):               # already on the device: pad there, no host round trip per image
        packed = torch.zeros(B, max_obj, 6, device=dev)
        for i, t in enumerate(targets):
            if t.size(0):
                packed[i, :t.size(0)] = t.detach().float()
    else:                                               # host targets (the reference's collate output): ONE stacked transfer
        packed = torch.zeros(B, max_obj, 6)
        for i, t in enumerate(targets):
            if t.size(0):
                packed[i, :t.size(0)] = t.detach().float().cpu()
        packed = packed.to(dev, non_blocking=True)
    nobj = nobj.to(dev, non_blocking=True)
    loc_t = torch.empty(B, P, 4, device=dev)
    conf_t = torch.empty(B, P, 2, device=dev)
    obj_u8 = torch.empty(B, P, dtype=torch.uint8, device=dev)
    bti = torch.empty(B, P, dtype=torch.int32, device=dev)
    ovl = torch.empty(B, P, device=dev) if want_overlap else None
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().ctx_match_encode(
            packed.data_ptr(), nobj.data_ptr(), max_obj, priors.data_ptr(), B, P, float(threshold),
            float(variances[0]), float(variances[1]), loc_t.data_ptr(), conf_t.data_ptr(), obj_u8.data_ptr(),
            bti.data_ptr(), ovl.data_ptr() if want_overlap else None, _lib.current_stream_ptr()), 'ctx_match_encode')
    out = (loc_t, conf_t, obj_u8 if obj_as_u8 else obj_u8.bool(), bti)
    return out + (ovl,) if want_overlap else out


def match(threshold, truths, priors, variances, labels, loc_t, conf_t, obj_t, idx, overlap=None):
    """Reference signature (box_utils.py:83): fills ``loc_t[idx]``, ``conf_t[idx]``, ``obj_t[idx]``
    (and ``overlap[idx]``) in place."""
    target = torch.cat([truths.float(), labels.float()], 1)
    res = match_batch(threshold, [target], priors, variances, want_overlap=overlap is not None)
    loc_t[idx] = res[0][0]
    conf_t[idx] = res[1][0]
    obj_t[idx] = res[2][0]
    if overlap is not None:
        overlap[idx] = res[4][0]


def hard_negative_rank(loss):
    """rank[b,p] = position of prior p in the descending sort of loss[b] — the result of the two
    sorts of multibox_loss_combined.py:91-93 (ties: lower index first)."""
    loss = _lib.require_cuda(loss, 'loss').float().contiguous()
    B, P = loss.shape
    L = _lib.lib()
    ws = torch.empty(L.ctx_rank_workspace_bytes(B, P), dtype=torch.uint8, device=loss.device)
    rank = torch.empty(B, P, dtype=torch.int32, device=loss.device)
    with torch.cuda.device(loss.device):
        _lib.check(L.ctx_hard_negative_rank(loss.data_ptr(), B, P, rank.data_ptr(), ws.data_ptr(), ws.numel(),
                                            _lib.current_stream_ptr()), 'ctx_hard_negative_rank')
    return rank

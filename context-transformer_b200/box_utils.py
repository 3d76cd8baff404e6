"""Box algebra of the detection path — mirror of the live part of reference ``utils/box_utils.py``.

``match`` (box_utils.py:83-132) keeps the reference's in-place signature but runs on the GPU
(``ctx_match_encode``); ``match_batch`` is the batched form the loss uses (one launch for the whole
batch instead of the Python loop of multibox_loss_combined.py:70-74).  ``decode`` (:184-202) goes
through the ``Detect`` kernel.  ``point_form`` / ``jaccard`` / ``encode`` are thin tensor
expressions kept for API completeness (they are not on the timed path).
"""
import torch

from . import _lib


def point_form(boxes):
    return torch.cat((boxes[:, :2] - boxes[:, 2:] / 2, boxes[:, :2] + boxes[:, 2:] / 2), 1)


def intersect(box_a, box_b):
    A, B = box_a.size(0), box_b.size(0)
    max_xy = torch.min(box_a[:, 2:].unsqueeze(1).expand(A, B, 2), box_b[:, 2:].unsqueeze(0).expand(A, B, 2))
    min_xy = torch.max(box_a[:, :2].unsqueeze(1).expand(A, B, 2), box_b[:, :2].unsqueeze(0).expand(A, B, 2))
    inter = torch.clamp((max_xy - min_xy), min=0)
    return inter[:, :, 0] * inter[:, :, 1]


def jaccard(box_a, box_b):
    inter = intersect(box_a, box_b)
    area_a = ((box_a[:, 2] - box_a[:, 0]) * (box_a[:, 3] - box_a[:, 1])).unsqueeze(1).expand_as(inter)
    area_b = ((box_b[:, 2] - box_b[:, 0]) * (box_b[:, 3] - box_b[:, 1])).unsqueeze(0).expand_as(inter)
    return inter / (area_a + area_b - inter)


def encode(matched, priors, variances):
    g_cxcy = (matched[:, :2] + matched[:, 2:]) / 2 - priors[:, :2]
    g_cxcy /= (variances[0] * priors[:, 2:])
    g_wh = (matched[:, 2:] - matched[:, :2]) / priors[:, 2:]
    g_wh = torch.log(g_wh) / variances[1]
    return torch.cat([g_cxcy, g_wh], 1)


def decode(loc, priors, variances):
    """[P,4] offsets + [P,4] priors -> [P,4] corner boxes (box_utils.py:184-202), on the GPU."""
    from .detection import Detect
    P = loc.size(0)
    dummy_conf = torch.zeros(1, P, 1, device=loc.device)
    dummy_obj = torch.zeros(1, P, 2, device=loc.device)
    boxes, _ = Detect(2, 0, {'variance': variances}).forward((loc.unsqueeze(0), dummy_conf, dummy_obj), priors)
    return boxes[0]


def match_batch(threshold, targets, priors, variances, want_overlap=False):
    """Batched ``match``: targets is a list of [n_i,6] tensors (x1,y1,x2,y2,label,weight).
    Returns loc_t[B,P,4] f32, conf_t[B,P,2] f32, obj_t[B,P] bool, best_truth_idx[B,P] int32
    (and best_truth_overlap[B,P] when asked)."""
    priors = _lib.require_cuda(priors, 'priors').float().contiguous()
    dev = priors.device
    B, P = len(targets), priors.size(0)
    max_obj = max([int(t.size(0)) for t in targets] + [1])
    packed = torch.zeros(B, max_obj, 6)
    nobj = torch.zeros(B, dtype=torch.int32)
    for i, t in enumerate(targets):
        n = int(t.size(0))
        nobj[i] = n
        if n:
            packed[i, :n] = t.detach().float().cpu()
    packed = packed.to(dev)
    nobj = nobj.to(dev)
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
    out = (loc_t, conf_t, obj_u8.bool(), bti)
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

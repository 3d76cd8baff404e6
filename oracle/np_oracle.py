"""TEST INFRASTRUCTURE ONLY — numpy fp32 restatement of the post-processing and target-matching
algorithms of the reference (never imported by the product package).

Every function cites the reference file:line it follows.  All arithmetic is float32 with the
reference's operation order so that integer outputs (kept indices, matched indices, ranks) are
bit-exact and float outputs agree to rounding of ``exp``/``log``.

Tie convention: the reference orders candidates with ``scores.argsort()[::-1]`` (numpy quicksort,
unstable — utils/nms/cpu_nms.pyx:25, gpu_nms.pyx:25-26), so the order of equal scores is
unspecified upstream.  The oracle (and the CUDA path) fix it as: score descending, then original
index ascending.  On tie-free inputs this is identical to the reference.
"""
import itertools
from math import sqrt

import numpy as np

F = np.float32


# --------------------------------------------------------------------------------------------
# PriorBox — layers/functions/prior_box.py:31-56
# --------------------------------------------------------------------------------------------
def prior_box(cfg):
    mean = []
    image_size = cfg['min_dim']
    for k, f in enumerate(cfg['feature_maps']):
        for i, j in itertools.product(range(f), repeat=2):
            f_k = image_size / cfg['steps'][k]
            cx = (j + 0.5) / f_k
            cy = (i + 0.5) / f_k
            s_k = cfg['min_sizes'][k] / image_size
            mean += [cx, cy, s_k, s_k]
            s_k_prime = sqrt(s_k * (cfg['max_sizes'][k] / image_size))
            mean += [cx, cy, s_k_prime, s_k_prime]
            for ar in cfg['aspect_ratios'][k]:
                mean += [cx, cy, s_k * sqrt(ar), s_k / sqrt(ar)]
                mean += [cx, cy, s_k / sqrt(ar), s_k * sqrt(ar)]
    out = np.asarray(mean, dtype=np.float64).astype(F).reshape(-1, 4)
    if cfg['clip']:
        out = np.clip(out, F(0), F(1))
    return out


# --------------------------------------------------------------------------------------------
# decode — utils/box_utils.py:184-202 ; Detect.forward — layers/functions/detection.py:18-55
# --------------------------------------------------------------------------------------------
def decode(loc, priors, variances=(0.1, 0.2)):
    loc = loc.astype(F, copy=False)
    priors = priors.astype(F, copy=False)
    v0, v1 = F(variances[0]), F(variances[1])
    cxcy = priors[:, :2] + (loc[:, :2] * v0) * priors[:, 2:]
    wh = priors[:, 2:] * np.exp(loc[:, 2:] * v1).astype(F)
    x1y1 = cxcy - wh / F(2)
    x2y2 = wh + x1y1
    return np.concatenate([x1y1, x2y2], axis=1).astype(F)


def detect(loc, conf, obj, priors, variances=(0.1, 0.2)):
    """(boxes[B,P,4], scores[B,P,1+C]) — scores = cat(obj[...,0], obj[...,1]*conf)."""
    B = loc.shape[0]
    boxes = np.stack([decode(loc[b], priors, variances) for b in range(B)])
    scores = np.concatenate([obj[..., 0:1], obj[..., 1:2] * conf], axis=2).astype(F)
    return boxes, scores


# --------------------------------------------------------------------------------------------
# hard NMS — utils/nms/cpu_nms.pyx:17-68 (">=" suppress), nms_kernel.cu:24-78 + :124-140 and
# py_cpu_nms.py:10-38 (">" suppress).  +1 pixel convention in both.
# --------------------------------------------------------------------------------------------
def sort_order(scores):
    """score descending, index ascending on ties."""
    return np.argsort(-scores.astype(F), kind='stable')


def iou_plus1(box, boxes):
    """fp32 IoU of one box against many with the +1 area convention (cpu_nms.pyx:24,56-63)."""
    one = F(1)
    area = (box[2] - box[0] + one) * (box[3] - box[1] + one)
    areas = (boxes[:, 2] - boxes[:, 0] + one) * (boxes[:, 3] - boxes[:, 1] + one)
    xx1 = np.maximum(box[0], boxes[:, 0])
    yy1 = np.maximum(box[1], boxes[:, 1])
    xx2 = np.minimum(box[2], boxes[:, 2])
    yy2 = np.minimum(box[3], boxes[:, 3])
    w = np.maximum(F(0), xx2 - xx1 + one)
    h = np.maximum(F(0), yy2 - yy1 + one)
    inter = (w * h).astype(F)
    with np.errstate(divide='ignore', invalid='ignore'):
        return (inter / (area + areas - inter)).astype(F)


def nms(dets, thresh, suppress_on_equal=False):
    """Greedy hard NMS.  dets[n,5] float32 (x1,y1,x2,y2,score).  Returns kept indices into dets in
    descending-score order.  suppress_on_equal=True is cpu_nms (``ovr >= thresh``), False is
    gpu_nms / py_cpu_nms (``ovr > thresh``)."""
    dets = np.ascontiguousarray(dets, dtype=F)
    n = dets.shape[0]
    if n == 0:
        return []
    thresh = F(thresh)
    order = sort_order(dets[:, 4])
    b = dets[order, :4]
    alive = np.ones(n, dtype=bool)
    keep = []
    for i in range(n):
        if not alive[i]:
            continue
        keep.append(int(order[i]))
        if i + 1 < n:
            ov = iou_plus1(b[i], b[i + 1:])
            sup = (ov >= thresh) if suppress_on_equal else (ov > thresh)
            alive[i + 1:] &= ~sup
    return keep


# --------------------------------------------------------------------------------------------
# soft-NMS — utils/nms/cpu_nms.pyx:70-163 (in-place; returns range(N_final))
# --------------------------------------------------------------------------------------------
def soft_nms(boxes, sigma=0.5, Nt=0.3, threshold=0.001, method=0):
    """Literal restatement (python loops; small n only).  Mutates a copy and returns
    (boxes_out[N_final,5], N_final).  method: 0 hard, 1 linear, 2 gaussian."""
    boxes = np.array(boxes, dtype=F, copy=True)
    N = boxes.shape[0]
    sigma, Nt, threshold = F(sigma), F(Nt), F(threshold)
    one = F(1)
    i = 0
    while i < N:
        maxscore = boxes[i, 4]
        maxpos = i
        t = boxes[i].copy()
        pos = i + 1
        while pos < N:
            if maxscore < boxes[pos, 4]:
                maxscore = boxes[pos, 4]
                maxpos = pos
            pos += 1
        boxes[i] = boxes[maxpos]
        boxes[maxpos] = t
        tx1, ty1, tx2, ty2, ts = boxes[i]
        pos = i + 1
        while pos < N:
            x1, y1, x2, y2, s = boxes[pos]
            # Cython emits the literal 1 as a double constant: these are double evaluations rounded
            # to float once on assignment (see oracle/c/nms_oracle.c).
            D = np.float64
            area = F((D(F(x2 - x1)) + 1.0) * (D(F(y2 - y1)) + 1.0))
            iw = F(D(F(min(tx2, x2) - max(tx1, x1))) + 1.0)
            if iw > 0:
                ih = F(D(F(min(ty2, y2) - max(ty1, y1))) + 1.0)
                if ih > 0:
                    ua = F(((D(F(tx2 - tx1)) + 1.0) * (D(F(ty2 - ty1)) + 1.0) + D(area)) - D(F(iw * ih)))
                    ov = F(F(iw * ih) / ua)
                    if method == 1:
                        weight = F(one - ov) if ov > Nt else one
                    elif method == 2:
                        weight = F(np.exp(np.float64(F(-(ov * ov) / sigma))))
                    else:
                        weight = F(0) if ov > Nt else one
                    boxes[pos, 4] = F(weight * boxes[pos, 4])
                    if boxes[pos, 4] < threshold:
                        boxes[pos] = boxes[N - 1]
                        N -= 1
                        pos -= 1
            pos += 1
        i += 1
    return boxes[:N].copy(), N


# --------------------------------------------------------------------------------------------
# test.py:133-161 post-processing: scale, per-class threshold + NMS, per-image top-k
# --------------------------------------------------------------------------------------------
def postprocess_image(boxes, scores, scale, thresh=0.01, nms_thresh=0.45, max_per_image=200,
                      suppress_on_equal=False, nms_fn=None):
    """boxes[P,4] fractional, scores[P,1+C].  Returns list over classes 1..C of dets[k,5] and the
    kept prior indices per class (same order)."""
    boxes = (boxes.astype(F) * np.asarray(scale, dtype=F)).astype(F)
    C1 = scores.shape[1]
    all_dets = [None] * C1
    all_idx = [None] * C1
    for j in range(1, C1):
        inds = np.where(scores[:, j] > F(thresh))[0]
        if len(inds) == 0:
            all_dets[j] = np.empty([0, 5], dtype=F)
            all_idx[j] = np.empty([0], dtype=np.int64)
            continue
        c_dets = np.hstack((boxes[inds], scores[inds, j][:, None])).astype(F, copy=False)
        keep = nms_fn(c_dets, nms_thresh) if nms_fn else nms(c_dets, nms_thresh, suppress_on_equal)
        all_dets[j] = c_dets[keep, :]
        all_idx[j] = inds[keep]
    if max_per_image > 0:
        image_scores = np.hstack([all_dets[j][:, -1] for j in range(1, C1)])
        if len(image_scores) > max_per_image:
            image_thresh = np.sort(image_scores)[-max_per_image]
            for j in range(1, C1):
                k = np.where(all_dets[j][:, -1] >= image_thresh)[0]
                all_dets[j] = all_dets[j][k, :]
                all_idx[j] = all_idx[j][k]
    return all_dets, all_idx


def records_from_dets(all_dets, all_idx=None):
    """Flatten per-class dets to records [k,6] = x1,y1,x2,y2,score,class in (class asc, score desc)
    order — the fixed-shape record format of the fused CUDA post-processing."""
    recs, idx = [], []
    for j in range(1, len(all_dets)):
        d = all_dets[j]
        if d is None or len(d) == 0:
            continue
        recs.append(np.hstack([d, np.full((len(d), 1), j, dtype=F)]))
        if all_idx is not None:
            idx.append(all_idx[j])
    if not recs:
        return np.empty((0, 6), dtype=F), np.empty((0,), dtype=np.int64)
    return np.vstack(recs).astype(F), (np.concatenate(idx) if idx else None)


# --------------------------------------------------------------------------------------------
# jaccard / match / encode — utils/box_utils.py:5-14, 29-68, 83-156
# --------------------------------------------------------------------------------------------
def point_form(priors):
    return np.concatenate([priors[:, :2] - priors[:, 2:] / F(2), priors[:, :2] + priors[:, 2:] / F(2)], 1).astype(F)


def jaccard(box_a, box_b):
    max_xy = np.minimum(box_a[:, None, 2:], box_b[None, :, 2:])
    min_xy = np.maximum(box_a[:, None, :2], box_b[None, :, :2])
    inter = np.clip(max_xy - min_xy, F(0), None)
    inter = (inter[:, :, 0] * inter[:, :, 1]).astype(F)
    area_a = ((box_a[:, 2] - box_a[:, 0]) * (box_a[:, 3] - box_a[:, 1]))[:, None]
    area_b = ((box_b[:, 2] - box_b[:, 0]) * (box_b[:, 3] - box_b[:, 1]))[None, :]
    union = area_a + area_b - inter
    with np.errstate(divide='ignore', invalid='ignore'):
        return (inter / union).astype(F)


def encode(matched, priors, variances=(0.1, 0.2)):
    v0, v1 = F(variances[0]), F(variances[1])
    g_cxcy = (matched[:, :2] + matched[:, 2:]) / F(2) - priors[:, :2]
    g_cxcy = g_cxcy / (v0 * priors[:, 2:])
    g_wh = (matched[:, 2:] - matched[:, :2]) / priors[:, 2:]
    with np.errstate(divide='ignore', invalid='ignore'):
        g_wh = np.log(g_wh).astype(F) / v1
    return np.concatenate([g_cxcy, g_wh], 1).astype(F)


def match(threshold, truths, priors, variances, labels):
    """Returns (loc[P,4] f32, conf[P,2] f32, obj[P] bool, best_truth_idx[P] int64,
    best_truth_overlap[P] f32 BEFORE the force-match fill).  First-index argmax on ties (what
    torch.max does on CPU); forced matches applied in GT order, last GT wins (box_utils.py:119-123)."""
    truths = truths.astype(F)
    labels = labels.astype(F)
    overlaps = jaccard(truths, point_form(priors))
    best_prior_idx = overlaps.argmax(1)
    best_truth_idx = overlaps.argmax(0)
    best_truth_overlap = overlaps.max(0)
    raw_overlap = best_truth_overlap.copy()
    best_truth_overlap[best_prior_idx] = F(2)
    for j in range(best_prior_idx.shape[0]):
        best_truth_idx[best_prior_idx[j]] = j
    matches = truths[best_truth_idx]
    conf = labels[best_truth_idx].copy()
    below = best_truth_overlap < F(threshold)
    conf[below, 0] = 0
    conf[below, 1] = 1
    loc = encode(matches, priors, variances)
    obj = conf[:, 0] != 0
    return loc, conf, obj, best_truth_idx, raw_overlap


# --------------------------------------------------------------------------------------------
# hard-negative ranking + loss — layers/modules/multibox_loss_combined.py:42-124
# --------------------------------------------------------------------------------------------
def _log_softmax(x):
    m = x.max(axis=-1, keepdims=True)
    return (x - m) - np.log(np.exp(x - m).sum(axis=-1, keepdims=True))


def hard_negative_rank(loss_obj):
    """rank[b,p] = position of p in the descending sort of loss_obj[b] (two sorts upstream,
    :91-93).  Ties: stable by index (torch.sort(stable=False) is unspecified on ties)."""
    order = np.argsort(-loss_obj, axis=1, kind='stable')
    rank = np.empty_like(order)
    B, P = loss_obj.shape
    rank[np.arange(B)[:, None], order] = np.arange(P)[None, :]
    return rank


def multibox_loss(loc_data, conf_data, obj_data, priors, targets, num_classes=21, threshold=0.5,
                  negpos_ratio=3, variances=(0.1, 0.2)):
    """float64-accumulated restatement; returns dict of python floats + the mining mask."""
    B, P = loc_data.shape[:2]
    loc_t = np.zeros((B, P, 4), F)
    conf_t = np.zeros((B, P, 2), F)
    obj_t = np.zeros((B, P), bool)
    for b in range(B):
        t = np.asarray(targets[b], dtype=F)
        loc_t[b], conf_t[b], obj_t[b], _, _ = match(threshold, t[:, :4], priors, variances, t[:, 4:6])
    pos = conf_t[:, :, 0] > 0
    num_pos = np.floor((conf_t[:, :, 1] * pos).sum(1, keepdims=True)).astype(np.int64)
    d = loc_data[pos] - loc_t[pos]
    ad = np.abs(d)
    sl1 = np.where(ad < 1, 0.5 * d * d, ad - 0.5)
    w_pos = conf_t[pos][:, 1]
    loss_l = float((sl1.sum(1) * w_pos).sum(dtype=np.float64))
    ls_obj = _log_softmax(obj_data.astype(F))
    ce_obj = -np.take_along_axis(ls_obj, obj_t.astype(np.int64)[..., None], axis=2)[..., 0]
    mine = ce_obj.copy()
    mine[obj_t] = 0
    rank = hard_negative_rank(mine)
    num_neg = np.minimum(negpos_ratio * num_pos, P - 1)
    neg = rank < num_neg
    mask = pos | neg
    weight = conf_t[mask][:, 1]
    loss_obj = float((ce_obj[mask] * weight).sum(dtype=np.float64))
    logit0 = obj_data[..., 0:1] + np.log(np.exp(conf_data).sum(axis=2, keepdims=True))
    logitk = obj_data[..., 1:2] + conf_data
    logit = np.concatenate([logit0, logitk], 2).astype(F)
    ls = _log_softmax(logit[mask])
    lab = conf_t[mask][:, 0].astype(np.int64)
    ce = -ls[np.arange(len(lab)), lab]
    loss_c = float((ce * weight).sum(dtype=np.float64))
    N = float(num_pos.sum())
    return {'loss_box_reg': loss_l / N, 'loss_cls': loss_c / N, 'loss_obj': loss_obj / N,
            'mask': mask, 'rank': rank, 'num_pos': num_pos}


# --------------------------------------------------------------------------------------------
# OBJ(Target) prototype initialisation — train.py:252-286 (init_reweight)
# --------------------------------------------------------------------------------------------
def init_reweight_prototypes(conf_batches, label_batches, num_classes=21, setting='transfer'):
    """conf_batches: list of [B,P,D] float32 (``model(data, init=True)``), label_batches: list of [B,P] float32
    (``conf_t[:, :, 0]`` from ``match``).  Follows train.py:268-282 literally: per class i the rows with label == i are
    concatenated over the batches in (batch, image, prior) order, each row divided by its L2 norm, averaged, then the
    mean divided by its own norm (:283-286).  Returns [n_classes_kept, D] float32 (NaN rows for classes without samples)."""
    n_fg = num_classes - 1
    cls = [np.zeros((0, conf_batches[0].shape[-1]), F) for _ in range(n_fg)]
    for conf, lab in zip(conf_batches, label_batches):
        for i in range(1, num_classes):
            cls[i - 1] = np.concatenate([cls[i - 1], conf[lab == i].astype(F)], 0)
    out = []
    for item in cls:
        with np.errstate(invalid='ignore', divide='ignore'):
            rows = item / np.sqrt((item * item).sum(1, keepdims=True, dtype=F)).astype(F)
            m = rows.mean(0, dtype=F) if len(rows) else np.full(item.shape[1], np.nan, F)
            out.append((m / np.sqrt((m * m).sum(dtype=F))).astype(F))
    if setting == 'incre':
        out = out[15:]
    return np.stack(out, 0)


# --------------------------------------------------------------------------------------------
# BaseTransform without the resize — data/data_augment.py:258-261 for an image of the network's size
# --------------------------------------------------------------------------------------------
def base_transform_same_size(img_u8, means=(104, 117, 123)):
    """img_u8 [H,W,3] uint8 -> [3,H,W] float32: ``img.astype(np.float32)``, ``img -= means``, ``transpose(2,0,1)``
    (cv2.resize to the image's own size is a plain copy, so these three statements are the whole transform)."""
    img = img_u8.astype(F)
    img -= np.asarray(means, F)
    return img.transpose(2, 0, 1)


# --------------------------------------------------------------------------------------------
# BaseTransform with the resize — data/data_augment.py:257-261: cv2.resize(img, (S, S), INTER_LINEAR) on uint8.
# Restatement of OpenCV's 8-bit bilinear algorithm (modules/imgproc/src/resize.cpp: cv::resize coordinate / coefficient set-up,
# HResizeLinear<uchar,int,short,INTER_RESIZE_COEF_SCALE = 2048>, VResizeLinear<uchar,int,short,FixedPtCast<int,uchar,22>>).
# PINNED: bit-exact against cv2.resize (opencv-python-headless 4.13, the build image's) on up- and down-scaling, odd sizes,
# 1x1 .. 1920x1080 sources (tests/test_oracle_golden.py, live) and against tests/golden/resize.npz, which oracle/gen_golden.py
# produces by running the REFERENCE's own BaseTransform class (data/data_augment.py:224-266, loaded from /root/reference).
# --------------------------------------------------------------------------------------------
def _cv_taps(dst, src, clamp_edges):
    """(s[dst], a0[dst], a1[dst]) of cv::resize for INTER_LINEAR, fixed point."""
    scale = 1.0 / (float(dst) / float(src))                       # double: inv_scale = dsize / ssize, scale = 1. / inv_scale
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)              # (float)((dx + 0.5) * scale_x - 0.5)
    s = np.floor(f).astype(np.int64)                              # cvFloor
    f = (f - s.astype(np.float32)).astype(np.float32)             # fx -= sx  (float)
    if clamp_edges:
        lo = s < 0
        f[lo], s[lo] = 0.0, 0
        hi = s >= src - 1
        f[hi], s[hi] = 0.0, src - 1
    a0 = np.rint((np.float32(1.0) - f) * np.float32(2048.0)).astype(np.int64)     # saturate_cast<short>(cvRound(..)): half to even
    a1 = np.rint(f * np.float32(2048.0)).astype(np.int64)
    return s, a0, a1


def cv_resize_linear_u8(img, size):
    """img [H,W,3] uint8 -> [size,size,3] uint8, cv2.resize(img, (size, size), interpolation=cv2.INTER_LINEAR)."""
    img = np.asarray(img, dtype=np.uint8)
    sh, sw = img.shape[:2]
    sx, a0, a1 = _cv_taps(size, sw, True)
    sy, b0, b1 = _cv_taps(size, sh, False)
    src = img.astype(np.int64)
    one = sx + 1 >= sw                                            # dx >= xmax: D = S[sx] * ONE
    sx1 = np.minimum(sx + 1, sw - 1)
    rows = np.where(one[None, :, None], src[:, sx] * 2048, src[:, sx] * a0[None, :, None] + src[:, sx1] * a1[None, :, None])   # [H, size, 3] int
    y0, y1 = np.clip(sy, 0, sh - 1), np.clip(sy + 1, 0, sh - 1)
    r0, r1 = rows[y0], rows[y1]
    v = (((b0[:, None, None] * (r0 >> 4)) >> 16) + ((b1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return (v & 255).astype(np.uint8)


def base_transform(img_u8, size, means=(104, 117, 123)):
    """data_augment.py:257-261: resize -> astype(float32) -> -= means -> transpose(2, 0, 1)."""
    img = cv_resize_linear_u8(img_u8, size).astype(F)
    img -= np.asarray(means, F)
    return img.transpose(2, 0, 1)

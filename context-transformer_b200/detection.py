"""``Detect`` and the fused post-processing of the inference path.

``Detect`` mirrors reference ``layers/functions/detection.py:6-55``: same constructor
``Detect(num_classes, bkg_label, cfg)``, same plain-method call ``detector.forward(predictions,
prior)`` returning ``(boxes[B,P,4], scores[B,P,num_classes])`` on the device of ``loc``, fresh
tensors each call, also kept on ``self.boxes`` / ``self.scores``.  Like the reference it does NOT
threshold or suppress.  The arithmetic (``decode`` of utils/box_utils.py:184-202, score combine
``cat(obj0, obj1*conf)``) runs in one CUDA kernel for the whole batch (``ctx_detect_forward``)
instead of a Python loop over images.

``DetectPost`` is the device-resident replacement of the numpy loop that follows ``Detect`` in
reference ``test.py:133-161`` (pixel scale, per-class ``score > thresh``, ``nms``, per-image
``max_per_image`` cut): one call, fixed-shape records out, no host round trips.
"""
import ctypes as C

import torch

from . import _lib

NMS_HARD, NMS_SOFT_LINEAR, NMS_SOFT_GAUSSIAN, NMS_SOFT_HARD = 0, 1, 2, 3


def _f32c(t, name):
    _lib.require_cuda(t, name)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class Detect(object):
    def __init__(self, num_classes, bkg_label, cfg):
        self.num_classes = num_classes
        self.background_label = bkg_label
        self.variance = cfg['variance']

    def forward(self, predictions, prior):
        loc, conf, obj = predictions
        loc_data = _f32c(loc.detach(), 'loc')
        conf_data = _f32c(conf.detach(), 'conf')
        obj_data = _f32c(obj.detach(), 'obj')
        prior_data = _f32c(prior.detach().to(loc_data.device), 'prior')
        num = loc_data.size(0)
        self.num_priors = prior_data.size(0)
        if conf_data.size(-1) != self.num_classes - 1:
            raise ValueError('Detect: conf has %d classes, expected num_classes-1 = %d'
                             % (conf_data.size(-1), self.num_classes - 1))
        self.boxes = torch.empty(num, self.num_priors, 4, device=loc_data.device)
        self.scores = torch.empty(num, self.num_priors, self.num_classes, device=loc_data.device)
        with torch.cuda.device(loc_data.device):
            _lib.check(_lib.lib().ctx_detect_forward(
                loc_data.data_ptr(), conf_data.data_ptr(), obj_data.data_ptr(), prior_data.data_ptr(),
                num, self.num_priors, self.num_classes - 1, float(self.variance[0]), float(self.variance[1]),
                self.boxes.data_ptr(), self.scores.data_ptr(), _lib.current_stream_ptr()), 'ctx_detect_forward')
        return self.boxes, self.scores

    __call__ = forward


class DetectPost(object):
    """Fused decode + score + per-class (soft-)NMS + top-k (reference test.py:133-161).

    forward(predictions, prior, scale) -> (records[B,max_out,6], counts[B], prior_idx[B,max_out])
      records rows are (x1, y1, x2, y2, score, class) in pixel units, ordered class ascending then
      score descending — exactly the rows ``all_boxes[j][i]`` of test.py would hold, flattened.
      ``counts[b]`` is the number of detections the reference keeps (ties at the top-k threshold
      can exceed ``max_per_image``; rows beyond ``max_out`` are dropped).
    ``scale`` is ``[w, h, w, h]`` (one for the batch) or ``[B,4]`` (test.py:123-124).
    ``suppress_on_equal`` selects the reference's CPU NMS convention (``ovr >= thresh``,
    cpu_nms.pyx:65) instead of the GPU/python one (``>``, nms_kernel.cu:71).
    """

    def __init__(self, num_classes, bkg_label, cfg, score_thresh=0.01, nms_thresh=0.45, max_per_image=200,
                 max_out=None, suppress_on_equal=False, nms_method=NMS_HARD, soft_sigma=0.5, soft_threshold=0.001):
        self.num_classes = num_classes
        self.background_label = bkg_label
        self.variance = cfg['variance']
        self.score_thresh = score_thresh
        self.nms_thresh = nms_thresh
        self.max_per_image = max_per_image
        self.max_out = max_out if max_out is not None else (max_per_image + 56 if max_per_image > 0 else 1024)
        self.suppress_on_equal = suppress_on_equal
        self.nms_method = nms_method
        self.soft_sigma = soft_sigma
        self.soft_threshold = soft_threshold
        self._ws = None

    def forward(self, predictions, prior, scale):
        loc, conf, obj = predictions
        loc = _f32c(loc.detach(), 'loc')
        conf = _f32c(conf.detach(), 'conf')
        obj = _f32c(obj.detach(), 'obj')
        dev = loc.device
        prior = _f32c(prior.detach().to(dev), 'prior')
        B, P = loc.size(0), prior.size(0)
        Cfg = self.num_classes - 1
        if conf.size(-1) != Cfg:
            raise ValueError('DetectPost: conf has %d classes, expected %d' % (conf.size(-1), Cfg))
        scale = torch.as_tensor(scale, dtype=torch.float32).to(dev).contiguous()
        per_image = int(scale.dim() == 2)
        if per_image and scale.size(0) != B:
            raise ValueError('DetectPost: scale must be [4] or [B,4]')
        p = _lib.CtxPostParams(B, P, Cfg, float(self.variance[0]), float(self.variance[1]), per_image,
                               float(self.score_thresh), float(self.nms_thresh), int(self.suppress_on_equal),
                               int(self.nms_method), float(self.soft_sigma), float(self.soft_threshold),
                               int(self.max_per_image), int(self.max_out))
        L = _lib.lib()
        need = L.ctx_postprocess_workspace_bytes(B, P, Cfg)
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        records = torch.zeros(B, self.max_out, 6, device=dev)
        counts = torch.zeros(B, dtype=torch.int32, device=dev)
        prior_idx = torch.full((B, self.max_out), -1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.ctx_detect_postprocess(
                loc.data_ptr(), conf.data_ptr(), obj.data_ptr(), prior.data_ptr(), scale.data_ptr(), C.byref(p),
                records.data_ptr(), counts.data_ptr(), prior_idx.data_ptr(),
                self._ws.data_ptr(), self._ws.numel(), _lib.current_stream_ptr()), 'ctx_detect_postprocess')
        return records, counts, prior_idx

    __call__ = forward


def records_to_all_boxes(records, counts, num_classes):
    """Host-side view of the records in the reference's result structure (test.py:107-108,154):
    ``all_boxes[j][i]`` = float32 ndarray [k,5] for class j, image i."""
    import numpy as np
    rec = records.cpu().numpy()
    cnt = counts.cpu().numpy()
    if int(cnt.max(initial=0)) > int(records.shape[1]):
        import warnings
        warnings.warn('DetectPost kept %d detections for an image but the record buffer holds %d rows: the excess (lowest classes '
                      'last) was dropped — raise max_out' % (int(cnt.max()), int(records.shape[1])))
    B = rec.shape[0]
    all_boxes = [[np.empty((0, 5), dtype=np.float32) for _ in range(B)] for _ in range(num_classes)]
    for i in range(B):
        r = rec[i, :min(int(cnt[i]), rec.shape[1])]
        cls = r[:, 5].astype(np.int64)
        for j in np.unique(cls):
            all_boxes[j][i] = r[cls == j, :5].copy()
    return all_boxes


class DetectionCollector(object):
    """The reference's evaluation hand-off (test.py:107-108, 150-154, 171-175): ``all_boxes[j][i]`` = float32 ndarray
    ``[k, 5]`` (x1, y1, x2, y2, score) for class j >= 1 and image i of the dataset (class 0 keeps the reference's empty
    lists), filled batch by batch from the fixed-shape records of ``DetectPost`` / ``shard.gather_records`` and pickled
    as ``detections.pkl`` so that ``dataset.evaluate_detections(all_boxes, save_folder)`` runs unchanged."""

    def __init__(self, num_images, num_classes):
        self.num_images, self.num_classes = int(num_images), int(num_classes)
        self.all_boxes = [[[] for _ in range(self.num_images)] for _ in range(self.num_classes)]

    def add(self, first_image, records, counts):
        """records [B, K, 6] / counts [B] of images first_image .. first_image + B - 1 (host or device tensors)."""
        import numpy as np
        batch = records_to_all_boxes(records, counts, self.num_classes)
        B = int(records.shape[0])
        if first_image < 0 or first_image + B > self.num_images:
            raise IndexError('images %d..%d outside the dataset of %d' % (first_image, first_image + B - 1, self.num_images))
        for j in range(1, self.num_classes):
            for i in range(B):
                self.all_boxes[j][first_image + i] = np.ascontiguousarray(batch[j][i], dtype=np.float32)
        return self

    def save(self, det_file):
        import pickle
        with open(det_file, 'wb') as f:
            pickle.dump(self.all_boxes, f, pickle.HIGHEST_PROTOCOL)
        return det_file

    # ---- the datasets' result files, written straight from the collected records --------------------------------------
    def write_voc_results(self, image_ids, class_names, filedir, filename='comp4_det_test_{:s}.txt'):
        """The per-class text files of ``VOCDetection._write_voc_results_file`` (data/voc0712.py:360-376): one line
        ``<image id> <score .3f> <x1+1 .1f> <y1+1 .1f> <x2+1 .1f> <y2+1 .1f>`` per detection, images in dataset order.
        ``image_ids``: the dataset's ``ids`` (``(root, id)`` pairs as upstream, or plain id strings); ``class_names``:
        ``VOC_CLASSES[split][:16 or 21]`` including ``'__background__'``.  Returns the list of files written."""
        import os
        os.makedirs(filedir, exist_ok=True)
        if len(image_ids) != self.num_images:
            raise ValueError('%d image ids for %d collected images' % (len(image_ids), self.num_images))
        written = []
        for cls_ind, cls in enumerate(class_names):
            if cls == '__background__':
                continue
            path = os.path.join(filedir, filename.format(cls))
            with open(path, 'wt') as f:
                for im_ind, index in enumerate(image_ids):
                    index = index[1] if isinstance(index, (tuple, list)) else index
                    dets = self.all_boxes[cls_ind][im_ind]
                    for k in range(len(dets)):
                        f.write('{:s} {:.3f} {:.1f} {:.1f} {:.1f} {:.1f}\n'.format(index, dets[k, -1], dets[k, 0] + 1, dets[k, 1] + 1,
                                                                                 dets[k, 2] + 1, dets[k, 3] + 1))
            written.append(path)
        return written

    def coco_results(self, img_ids, class_names, class_to_coco_cat_id):
        """The result list of ``COCODetection._write_coco_results_file`` (data/coco.py:242-274): per class (1-based, in
        ``class_names`` order) and image, ``{'image_id', 'category_id', 'bbox': [x, y, w + 1, h + 1], 'score'}`` in float64."""
        import numpy as np
        if len(img_ids) != self.num_images:
            raise ValueError('%d image ids for %d collected images' % (len(img_ids), self.num_images))
        results = []
        for cls_ind, cls in enumerate(class_names, 1):
            cat_id = class_to_coco_cat_id[cls]
            for im_ind, index in enumerate(img_ids):
                dets = np.asarray(self.all_boxes[cls_ind][im_ind], dtype=np.float64)
                if dets.size == 0:
                    continue
                xs, ys = dets[:, 0], dets[:, 1]
                ws, hs = dets[:, 2] - xs + 1, dets[:, 3] - ys + 1
                results.extend([{'image_id': index, 'category_id': cat_id, 'bbox': [xs[k], ys[k], ws[k], hs[k]], 'score': dets[k, -1]}
                                for k in range(dets.shape[0])])
        return results

    def write_coco_results(self, res_file, img_ids, class_names, class_to_coco_cat_id):
        import json
        with open(res_file, 'w') as fid:
            json.dump(self.coco_results(img_ids, class_names, class_to_coco_cat_id), fid)
            fid.flush()
        return res_file


class BaseTransform(object):
    """Mirror of reference ``data/data_augment.py:224-266`` for the inference path: ``transform(img)`` -> fp32 ``[3,S,S]``
    on the device.  The whole transform runs on the GPU: the 8-bit bilinear resize (``cv2.resize(.., INTER_LINEAR)``: OpenCV's
    fixed-point kernel restated operation for operation in ``ctx_base_transform_resize``), the mean subtraction and the
    HWC -> CHW change; a uint8 image crosses PCIe instead of an fp32 one.  Batches of images that already have the network's
    size: ``transform.batch(imgs_u8[B,S,S,3])`` -> ``[B,3,S,S]``, or pass the uint8 batch straight to ``net(...)``; lists of
    images of any sizes: ``transform.batch([img0, img1, ...])``."""

    def __init__(self, resize, rgb_means, swap=(2, 0, 1), device='cuda'):
        if tuple(swap) != (2, 0, 1):
            raise ValueError('BaseTransform: only the reference swap (2, 0, 1) is supported')
        self.resize, self.means, self.swap, self.device = int(resize), tuple(float(m) for m in rgb_means), tuple(swap), device

    def _one(self, img, out):
        import ctypes as C
        import torch
        t = torch.as_tensor(img)
        if t.dtype != torch.uint8 or t.dim() != 3 or t.size(2) != 3:
            raise ValueError('BaseTransform expects a uint8 [H,W,3] image (cv2.imread layout)')
        t = _lib.require_cuda(t.to(self.device, non_blocking=True).contiguous(), 'img')
        means = (C.c_float * 3)(*self.means)
        with torch.cuda.device(t.device):
            _lib.check(_lib.lib().ctx_base_transform_resize(t.data_ptr(), t.size(0), t.size(1), out.data_ptr(), self.resize, means,
                                                            _lib.current_stream_ptr()), 'ctx_base_transform_resize')

    def batch(self, imgs):
        import ctypes as C
        import torch
        if isinstance(imgs, (list, tuple)):                       # images of any sizes: one resize launch each
            out = torch.empty(len(imgs), 3, self.resize, self.resize, device=self.device)
            for i, im in enumerate(imgs):
                self._one(im, out[i])
            return out
        t = torch.as_tensor(imgs)
        if t.dtype != torch.uint8 or t.dim() != 4 or t.size(3) != 3:
            raise ValueError('BaseTransform.batch expects uint8 [B,H,W,3] or a list of uint8 [H,W,3] images')
        if t.size(1) != self.resize or t.size(2) != self.resize:
            return self.batch([t[i] for i in range(t.size(0))])
        t = t.to(self.device).contiguous()
        _lib.require_cuda(t, 'imgs')
        out = torch.empty(t.size(0), 3, self.resize, self.resize, device=t.device)
        means = (C.c_float * 3)(*self.means)
        with torch.cuda.device(t.device):
            _lib.check(_lib.lib().ctx_base_transform(t.data_ptr(), out.data_ptr(), t.size(0), self.resize, self.resize, means,
                                                     _lib.current_stream_ptr()), 'ctx_base_transform')
        return out

    def __call__(self, img, target=None):
        import torch
        if target is not None:
            raise NotImplementedError('BaseTransform with targets is the training-time path (host side, data_augment.py:249-256)')
        out = torch.empty(3, self.resize, self.resize, device=self.device)
        self._one(img, out)
        return out

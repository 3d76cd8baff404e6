"""TEST INFRASTRUCTURE ONLY — generate tests/golden/*.npz by running the REAL reference
(/root/reference, imported by oracle/ref_import.py) on the seeded synthetic inputs of
oracle/synth.py.  Run in the build container only:  python -m oracle.gen_golden

The reference ships no tests or golden vectors, so these files are what pins the
oracle (and through it the CUDA path) to the reference's behaviour.  Large tensors are stored
sub-sampled (every ``ROW_STRIDE``-th prior) together with float64 checksums of the full tensor.
"""
import os
import sys
import types

import numpy as np
import torch

from . import ref_import, synth

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
ROW_STRIDE = 5

NET_CASES = [  # tag, method, phase, setting, size, num_classes, batch
    ('ours_transfer_300', 'ours', 2, 'transfer', 300, 60, 2),
    ('ours_incre_300', 'ours', 2, 'incre', 300, 15, 1),
    ('ft_300', 'ft', 2, 'transfer', 300, 20, 1),
    ('ft_512', 'ft', 2, 'transfer', 512, 20, 1),
]


def checksum(a):
    a = np.asarray(a, dtype=np.float64)
    return np.array([a.sum(), np.abs(a).sum(), (a * a).sum()])


def gen_priors(r):
    out = {}
    for name in ('VOC_300', 'VOC_512', 'COCO_300', 'COCO_512'):
        p = r.PriorBox(getattr(r.cfg, name)).forward().numpy()
        out[name] = p
    np.savez_compressed(os.path.join(GOLD, 'priors.npz'), **out)


def gen_net(r):
    torch.set_num_threads(os.cpu_count() or 1)
    for tag, method, phase, setting, size, ncls, batch in NET_CASES:
        args = types.SimpleNamespace(method=method, phase=phase, setting=setting)
        torch.manual_seed(0)
        net = r.build_net(args, size, ncls)
        net.load_state_dict(synth.seeded_state(net.state_dict(), seed=0))
        net.eval()
        net.device = 'cpu'
        x = synth.seeded_input(batch, size, seed=0)
        with torch.no_grad():
            loc, conf, obj = net(x)
            conf_init = net(x, init=True)
            net.train()
            loc_tr, conf_tr, obj_tr = net(x) if batch > 1 else (None, None, None)   # BN needs batch > 1 on 1x1 maps
        d = dict(keys=np.array(list(net.state_dict().keys())),
                 shapes=np.array([str(tuple(v.shape)) for v in net.state_dict().values()]),
                 loc=loc.numpy()[:, ::ROW_STRIDE], conf=conf.numpy()[:, ::ROW_STRIDE], obj=obj.numpy()[:, ::ROW_STRIDE],
                 conf_init=conf_init.numpy()[:, ::ROW_STRIDE],
                 loc_sum=checksum(loc), conf_sum=checksum(conf), obj_sum=checksum(obj),
                 conf_argmax=conf.argmax(-1).numpy().astype(np.int16))
        np.savez_compressed(os.path.join(GOLD, 'net_%s.npz' % tag), **d)
        print(tag, loc.shape, conf.shape, float(conf.max()), flush=True)


def gen_autocast(r):
    """The REFERENCE module under torch.autocast('cpu', bfloat16) on the same seeded state / input as net_ours_transfer_300:
    what 16-bit arithmetic costs the reference itself (SURVEY.md App. B measured 99.38 % argmax agreement on default-init
    weights; this pins the figure on the seeded state the GPU tests use), the yardstick for the bf16 / fp16 engine modes."""
    tag, method, phase, setting, size, ncls, batch = NET_CASES[0]
    args = types.SimpleNamespace(method=method, phase=phase, setting=setting)
    torch.manual_seed(0)
    net = r.build_net(args, size, ncls)
    net.load_state_dict(synth.seeded_state(net.state_dict(), seed=0))
    net.eval()
    net.device = 'cpu'
    x = synth.seeded_input(batch, size, seed=0)
    with torch.no_grad():
        ref = net(x)
        with torch.autocast('cpu', dtype=torch.bfloat16):
            ac = [t.float() for t in net(x)]
    d = dict(loc=ac[0].numpy()[:, ::ROW_STRIDE], conf=ac[1].numpy()[:, ::ROW_STRIDE], obj=ac[2].numpy()[:, ::ROW_STRIDE],
             conf_argmax=ac[1].argmax(-1).numpy().astype(np.int16),
             max_abs=np.array([float((a - b).abs().max()) for a, b in zip(ac, ref)]),
             mean_abs=np.array([float((a - b).abs().mean()) for a, b in zip(ac, ref)]),
             argmax_agreement=np.array(float((ac[1].argmax(-1) == ref[1].argmax(-1)).float().mean())))
    np.savez_compressed(os.path.join(GOLD, 'net_%s_autocast_bf16.npz' % tag), **d)
    print('autocast bf16 vs fp32 (reference, seeded state): max', d['max_abs'], 'mean', d['mean_abs'], 'argmax agreement', d['argmax_agreement'])


RESIZE_CASES = [((375, 500), 300, 0), ((500, 333), 300, 1), ((281, 500), 512, 2), ((64, 48), 300, 3), ((600, 600), 300, 4), ((300, 300), 300, 5),
                ((1, 7), 300, 6), ((720, 1280), 512, 7)]          # (source H, W), network size, seed


def resize_image(hw, seed):
    """Seeded uint8 test image: smooth gradients + noise (exercises both the interpolation and the rounding)."""
    g = np.random.default_rng(1000 + seed)
    h, w = hw
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(xx * 255.0 / max(w - 1, 1)), (yy * 255.0 / max(h - 1, 1)), ((xx + yy) % 256)], -1)
    return np.clip(base + g.normal(0, 40, (h, w, 3)), 0, 255).astype(np.uint8)


def gen_resize(r):
    """The reference's own BaseTransform (data/data_augment.py:224-266; it calls cv2.resize(.., INTER_LINEAR)) on seeded images."""
    import importlib.util
    saved = list(sys.path)
    sys.path.insert(0, ref_import.REF_ROOT)
    try:
        spec = importlib.util.spec_from_file_location('ref_data_augment', os.path.join(ref_import.REF_ROOT, 'data', 'data_augment.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path[:] = saved
    out = {}
    for hw, size, seed in RESIZE_CASES:
        t = mod.BaseTransform(size, (104, 117, 123))(resize_image(hw, seed)).numpy()       # [3, S, S] float32
        key = '%dx%d_%d' % (hw[0], hw[1], size)
        out['rows_' + key] = t[:, ::17, :].copy()                                              # every 17th row, all columns
        out['sum_' + key] = checksum(t)
    np.savez_compressed(os.path.join(GOLD, 'resize.npz'), **out)
    print('resize golden:', sorted(k for k in out if k.startswith('sum_')))


def _reference_functions(path, names, namespace):
    """Compile the named function definitions straight out of a reference source file (read at generation time, never
    stored in this repo) into ``namespace`` — for reference modules that cannot be imported whole here (``data/`` pulls in
    matplotlib and the compiled COCO mask extension)."""
    import ast
    tree = ast.parse(open(path).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names:
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(mod, path, 'exec'), namespace)
    return namespace


class _Dets(np.ndarray):
    """ndarray whose ``== []`` is False instead of a broadcast error: the reference's writers test ``if dets == []`` on
    arrays (numpy 1.x semantics)."""
    def __eq__(self, other):
        if isinstance(other, list):
            return False
        return np.ndarray.__eq__(self, other)


def writer_inputs(golden_post):
    """all_boxes[class][image] in the reference's structure, from the post-processing golden (2 images, 21 classes)."""
    num_images, num_classes = 2, 21
    all_boxes = [[[] for _ in range(num_images)] for _ in range(num_classes)]
    for b in range(num_images):
        rec = golden_post['records_gt_%d' % b]
        for j in range(1, num_classes):
            rows = rec[rec[:, 5] == j][:, :5].astype(np.float32)
            if len(rows):
                all_boxes[j][b] = rows
    ids = [('/data/VOCdevkit/VOC2007', '000001'), ('/data/VOCdevkit/VOC2007', '004242')]
    return all_boxes, ids


def gen_writers(r):
    """VOCDetection._write_voc_results_file (data/voc0712.py:360-376) and COCODetection._coco_results_one_category
    (data/coco.py:242-259) EXECUTED from the reference's source on the detections of the post-processing golden."""
    import tempfile
    import types as _t
    post = np.load(os.path.join(GOLD, 'post_voc300.npz'), allow_pickle=True)
    all_boxes, ids = writer_inputs(post)
    wrapped = [[(b.view(_Dets) if isinstance(b, np.ndarray) else b) for b in cls] for cls in all_boxes]
    ns = {'os': os, 'np': np, 'print': lambda *a, **k: None}
    src = open(os.path.join(ref_import.REF_ROOT, 'data', 'voc0712.py')).read()
    vc = {}
    exec(src[src.index('VOC_CLASSES = dict()'):src.index('# for making bounding boxes pretty')], vc)
    ns['VOC_CLASSES'] = vc['VOC_CLASSES']
    _reference_functions(os.path.join(ref_import.REF_ROOT, 'data', 'voc0712.py'), ('_write_voc_results_file', '_get_voc_results_file_template'), ns)
    tmp = tempfile.mkdtemp(prefix='ctx_voc_')
    fake = _t.SimpleNamespace(split=0, phase=2, ids=ids, root=tmp, _year='2007')
    fake._get_voc_results_file_template = lambda: ns['_get_voc_results_file_template'](fake)
    ns['_write_voc_results_file'](fake, wrapped)
    out = {}
    d = os.path.join(tmp, 'results', 'VOC2007', 'Main')
    for fn in sorted(os.listdir(d)):
        out['voc_' + fn] = np.array(open(os.path.join(d, fn)).read())
    cns = {'np': np}
    np.float = float                       # removed in numpy 1.24; the reference calls dets.astype(np.float)
    try:
        _reference_functions(os.path.join(ref_import.REF_ROOT, 'data', 'coco.py'), ('_coco_results_one_category',), cns)
        fake_c = _t.SimpleNamespace(img_ids=[139, 285])
        res = []
        for cls_ind in range(1, 21):
            res.extend(cns['_coco_results_one_category'](fake_c, wrapped[cls_ind], 100 + cls_ind))
    finally:
        del np.float
    out['coco_image_id'] = np.array([x['image_id'] for x in res])
    out['coco_category_id'] = np.array([x['category_id'] for x in res])
    out['coco_bbox'] = np.array([x['bbox'] for x in res], dtype=np.float64)
    out['coco_score'] = np.array([x['score'] for x in res], dtype=np.float64)
    np.savez_compressed(os.path.join(GOLD, 'writers.npz'), **out)
    print('writers golden: %d VOC files, %d COCO results' % (len([k for k in out if k.startswith('voc_')]), len(res)))


def gen_post(r):
    """Detect.forward + the numpy loop of test.py:133-161 with the reference's own NMS routines."""
    priors = r.PriorBox(r.cfg.VOC_300).forward()
    B, P, C = 2, priors.size(0), 20
    loc, conf, obj = synth.calibrated_heads(B, P, C, seed=0)
    det = r.Detect(21, 0, r.cfg.VOC_300)
    boxes, scores = det.forward((loc, conf, obj), priors)
    scale = np.array([500., 375., 500., 375.], dtype=np.float32)
    out = dict(boxes=boxes.numpy()[:, ::ROW_STRIDE], scores=scores.numpy()[:, ::ROW_STRIDE],
               boxes_sum=checksum(boxes), scores_sum=checksum(scores), scale=scale)
    for conv, fn in (('gt', r.py_cpu_nms), ('ge', r.cpu_nms)):
        if fn is None:
            continue
        for b in range(B):
            bx = (boxes[b] * torch.from_numpy(scale)).numpy()
            sc = scores[b].numpy()
            all_dets = [np.empty((0, 5), np.float32)] * 21
            all_idx = [np.empty((0,), np.int64)] * 21
            for j in range(1, 21):
                inds = np.where(sc[:, j] > 0.01)[0]
                if len(inds) == 0:
                    continue
                c_dets = np.hstack((bx[inds], sc[inds, j][:, None])).astype(np.float32, copy=False)
                keep = np.asarray(fn(c_dets, 0.45), dtype=np.int64)
                all_dets[j] = c_dets[keep]
                all_idx[j] = inds[keep]
            image_scores = np.hstack([all_dets[j][:, -1] for j in range(1, 21)])
            if len(image_scores) > 200:
                th = np.sort(image_scores)[-200]
                for j in range(1, 21):
                    k = np.where(all_dets[j][:, -1] >= th)[0]
                    all_dets[j] = all_dets[j][k]
                    all_idx[j] = all_idx[j][k]
            rec = np.vstack([np.hstack([all_dets[j], np.full((len(all_dets[j]), 1), j, np.float32)]) for j in range(1, 21)])
            out['records_%s_%d' % (conv, b)] = rec.astype(np.float32)
            out['prior_idx_%s_%d' % (conv, b)] = np.concatenate([all_idx[j] for j in range(1, 21)]).astype(np.int32)
    np.savez_compressed(os.path.join(GOLD, 'post_voc300.npz'), **out)


def gen_nms(r):
    out = {}
    for n, seed in ((0, 0), (1, 1), (63, 2), (64, 3), (65, 4), (300, 5), (2000, 6)):
        d = synth.random_dets(n, seed=seed)
        out['keep_gt_%d' % n] = np.asarray(r.py_cpu_nms(d, 0.45) if n else [], dtype=np.int32)
        if r.cpu_nms is not None:
            out['keep_ge_%d' % n] = np.asarray(r.cpu_nms(d, 0.45) if n else [], dtype=np.int32)
    if r.cpu_soft_nms is not None:
        for method in (0, 1, 2):
            for n, seed in ((1, 1), (65, 4), (300, 5)):
                d = synth.random_dets(n, seed=seed).copy()
                keep = r.cpu_soft_nms(d, sigma=0.5, Nt=0.3, threshold=0.001, method=method)
                out['soft_m%d_%d' % (method, n)] = d[:len(keep)].copy()
    np.savez_compressed(os.path.join(GOLD, 'nms.npz'), **out)


def gen_match_loss(r):
    priors = r.PriorBox(r.cfg.VOC_300).forward()
    P = priors.size(0)
    B = 4
    targets = synth.synthetic_targets(B, seed=0)
    targets[1][0, 4] = -1.0           # an "ignore" label (counts as object, box_utils.py:130)
    targets[2][:, 5] = 0.7            # mixup weights
    loc_t = torch.zeros(B, P, 4)
    conf_t = torch.zeros(B, P, 2)
    obj_t = torch.zeros(B, P, dtype=torch.bool)
    overlap = torch.zeros(B, P)
    for i in range(B):
        r.box_utils.match(0.5, targets[i][:, :4], priors, [0.1, 0.2], targets[i][:, 4:6], loc_t, conf_t, obj_t, i, overlap)
    g = synth._gen(0, 'losspred')
    loc_p = torch.randn(B, P, 4, generator=g).requires_grad_()
    conf_p = torch.randn(B, P, 20, generator=g).requires_grad_()
    obj_p = torch.randn(B, P, 2, generator=g).requires_grad_()
    crit = r.MultiBoxLoss_combined(21, 0.5, True, 0, True, 3, 0.5, False)
    losses = crit((loc_p, conf_p, obj_p), priors, targets)
    # gradients of the reference module's own autograd graph; distinct weights per term so that each gradient path is pinned
    (1.0 * losses['loss_box_reg'] + 2.0 * losses['loss_cls'] + 3.0 * losses['loss_obj']).backward()
    pos = conf_t[:, :, 0] != 0
    out = dict(loc_t_pos=loc_t[pos].numpy(), pos_index=pos.nonzero().numpy().astype(np.int32),
               conf_t=conf_t.numpy(), obj_t=obj_t.numpy(), overlap=overlap.numpy()[:, ::ROW_STRIDE],
               loc_t_sum=checksum(loc_t), targets=np.array([t.numpy() for t in targets], dtype=object),
               loss=np.array([float(losses['loss_box_reg']), float(losses['loss_cls']), float(losses['loss_obj'])]),
               grad_weights=np.array([1.0, 2.0, 3.0]), grad_loc=loc_p.grad.numpy(), grad_conf=conf_p.grad.numpy(), grad_obj=obj_p.grad.numpy())
    np.savez_compressed(os.path.join(GOLD, 'match_loss.npz'), **out)


def reweight_inputs(P, seed=0, batches=3, B=4, D=60):
    """Seeded inputs of the init_reweight golden: `batches` batches of raw conf features [B,P,D] and config-5 targets
    whose labels cover only part of the 20 classes (classes without a sample must come out NaN, as upstream)."""
    feats, targets = [], []
    for it in range(batches):
        g = synth._gen(seed + it, 'reweight')
        feats.append(torch.randn(B, P, D, generator=g) * 3.0 + 0.5)
        t = synth.synthetic_targets(B, seed=100 + seed + it, num_classes=12)
        targets.append(t)
    targets[1][0][0, 4] = -1.0                       # an "ignore" label: never a class sample (train.py:276 compares == i)
    return feats, targets


def gen_reweight(r):
    """train.py:252-286 executed literally on CPU tensors with the reference's own match() (train.py itself needs the
    dataset / logger stack and is not importable; the statements below are its lines :257-286 with the model call
    replaced by the seeded feature tensor)."""
    priors = r.PriorBox(r.cfg.VOC_300).forward()
    P = priors.size(0)
    num_classes = 21
    feats, targets_all = reweight_inputs(P)
    cls_list = [torch.empty(0) for _ in range(num_classes - 1)]
    labels_out = []
    for conf_data, targets in zip(feats, targets_all):
        num = conf_data.size(0)
        loc_t = torch.Tensor(num, P, 4)
        conf_t = torch.Tensor(num, P, 2)
        obj_t = torch.BoolTensor(num, P)
        for idx in range(num):
            truths = targets[idx][:, :-2].data
            labels = targets[idx][:, -2:].data
            r.box_utils.match(0.5, truths, priors.data, [0.1, 0.2], labels, loc_t, conf_t, obj_t, idx)
        conf_data_list = [conf_data[conf_t[:, :, 0] == i] for i in range(1, num_classes)]
        cls_list = [torch.cat((cls_list[i], conf_data_list[i]), 0) for i in range(num_classes - 1)]
        labels_out.append(conf_t[:, :, 0].numpy().copy())
    counts = np.array([len(c) for c in cls_list], np.int32)
    cls_list = [(item / item.norm(dim=1, keepdim=True)).mean(0) for item in cls_list]
    weight = torch.stack([item / item.norm() for item in cls_list], 0)
    np.savez_compressed(os.path.join(GOLD, 'reweight.npz'), weight=weight.numpy(), weight_incre=weight[15:].numpy(), counts=counts,
                        labels_checksum=np.array([checksum(torch.from_numpy(l)) for l in labels_out]))


def main():
    os.makedirs(GOLD, exist_ok=True)
    r = ref_import.load()
    which = sys.argv[1:] or ['priors', 'nms', 'post', 'match', 'net', 'reweight', 'autocast', 'resize', 'writers']
    if 'priors' in which:
        gen_priors(r)
    if 'nms' in which:
        gen_nms(r)
    if 'post' in which:
        gen_post(r)
    if 'match' in which:
        gen_match_loss(r)
    if 'reweight' in which:
        gen_reweight(r)
    if 'net' in which:
        gen_net(r)
    if 'autocast' in which:
        gen_autocast(r)
    if 'resize' in which:
        gen_resize(r)
    if 'writers' in which:
        gen_writers(r)


if __name__ == '__main__':
    main()

"""CPU: the oracle (numpy / C / torch restatements) against the golden vectors produced by the
real reference (oracle/gen_golden.py), and — when /root/reference is present — against the live
reference itself."""
import types

import numpy as np
import pytest
import torch

from oracle import c_oracle, np_oracle, ref_import, synth, torch_net
from oracle.gen_golden import NET_CASES, ROW_STRIDE, checksum

import context_transformer_b200 as ctx

CFGS = {'VOC_300': ctx.VOC_300, 'VOC_512': ctx.VOC_512, 'COCO_300': ctx.COCO_300, 'COCO_512': ctx.COCO_512}


@pytest.mark.parametrize('name', sorted(CFGS))
def test_prior_box_bit_exact(golden, name):
    g = golden('priors.npz')[name]
    assert np.array_equal(np_oracle.prior_box(CFGS[name]), g)
    assert g.shape[0] == ctx.num_priors(CFGS[name])


@pytest.mark.parametrize('n', [0, 1, 63, 64, 65, 300, 2000])
def test_hard_nms_matches_reference(golden, n):
    g = golden('nms.npz')
    seed = {0: 0, 1: 1, 63: 2, 64: 3, 65: 4, 300: 5, 2000: 6}[n]
    d = synth.random_dets(n, seed=seed)
    for conv, on_equal in (('gt', False), ('ge', True)):
        want = g['keep_%s_%d' % (conv, n)].tolist()
        assert np_oracle.nms(d, 0.45, on_equal) == want
        assert c_oracle.cpu_nms(d, 0.45, on_equal) == want


@pytest.mark.parametrize('method', [0, 1, 2])
@pytest.mark.parametrize('n', [1, 65, 300])
def test_soft_nms_matches_reference(golden, method, n):
    g = golden('nms.npz')['soft_m%d_%d' % (method, n)]
    seed = {1: 1, 65: 4, 300: 5}[n]
    d = synth.random_dets(n, seed=seed)
    out_c, n_c = c_oracle.cpu_soft_nms(d, 0.5, 0.3, 0.001, method)
    assert n_c == len(g) and np.array_equal(out_c, g)
    if n <= 65:
        out_np, n_np = np_oracle.soft_nms(d, 0.5, 0.3, 0.001, method)
        assert n_np == len(g) and np.array_equal(out_np, g)


def test_detect_and_postprocess_match_reference(golden):
    g = golden('post_voc300.npz')
    priors = np_oracle.prior_box(ctx.VOC_300)
    loc, conf, obj = [t.numpy() for t in synth.calibrated_heads(2, priors.shape[0], 20, seed=0)]
    boxes, scores = np_oracle.detect(loc, conf, obj, priors)
    assert np.allclose(boxes[:, ::ROW_STRIDE], g['boxes'], rtol=0, atol=2e-6)
    assert np.array_equal(scores[:, ::ROW_STRIDE], g['scores'])
    for conv, on_equal in (('gt', False), ('ge', True)):
        for b in range(2):
            dets, idx = np_oracle.postprocess_image(boxes[b], scores[b], g['scale'], suppress_on_equal=on_equal)
            rec, pidx = np_oracle.records_from_dets(dets, idx)
            assert np.array_equal(pidx, g['prior_idx_%s_%d' % (conv, b)])
            assert np.allclose(rec, g['records_%s_%d' % (conv, b)], rtol=0, atol=1e-3)
            assert np.array_equal(rec[:, 4:], g['records_%s_%d' % (conv, b)][:, 4:])
            assert 150 < len(rec) <= 200


def test_match_and_loss_match_reference(golden):
    g = golden('match_loss.npz')
    priors = np_oracle.prior_box(ctx.VOC_300)
    targets = [np.asarray(t, dtype=np.float32) for t in g['targets']]
    B, P = len(targets), priors.shape[0]
    loc_t = np.zeros((B, P, 4), np.float32)
    conf_t = np.zeros((B, P, 2), np.float32)
    obj_t = np.zeros((B, P), bool)
    ovl = np.zeros((B, P), np.float32)
    for i, t in enumerate(targets):
        loc_t[i], conf_t[i], obj_t[i], _, ovl[i] = np_oracle.match(0.5, t[:, :4], priors, (0.1, 0.2), t[:, 4:6])
    assert np.array_equal(conf_t, g['conf_t'])
    assert np.array_equal(obj_t, g['obj_t'])
    assert np.allclose(ovl[:, ::ROW_STRIDE], g['overlap'], rtol=0, atol=1e-6)
    pos = conf_t[:, :, 0] != 0
    assert np.array_equal(np.argwhere(pos).astype(np.int32), g['pos_index'])
    assert np.allclose(loc_t[pos], g['loc_t_pos'], rtol=0, atol=2e-5)
    gen = synth._gen(0, 'losspred')
    loc_p = torch.randn(B, P, 4, generator=gen).numpy()
    conf_p = torch.randn(B, P, 20, generator=gen).numpy()
    obj_p = torch.randn(B, P, 2, generator=gen).numpy()
    out = np_oracle.multibox_loss(loc_p, conf_p, obj_p, priors, targets)
    got = np.array([out['loss_box_reg'], out['loss_cls'], out['loss_obj']])
    assert np.allclose(got, g['loss'], rtol=2e-5)


@pytest.mark.parametrize('case', NET_CASES, ids=[c[0] for c in NET_CASES])
def test_torch_net_matches_reference(golden, case):
    tag, method, phase, setting, size, ncls, batch = case
    g = golden('net_%s.npz' % tag)
    args = types.SimpleNamespace(method=method, phase=phase, setting=setting)
    net = ctx.build_net(args, size, ncls)
    assert list(net.state_dict().keys()) == g['keys'].tolist()
    assert [str(tuple(v.shape)) for v in net.state_dict().values()] == g['shapes'].tolist()
    sd = synth.seeded_state(net.state_dict(), seed=0)
    x = synth.seeded_input(batch, size, seed=0)
    with torch.no_grad():
        loc, conf, obj = torch_net.forward(sd, x, size, ncls, method, phase, setting)
    # fp32 restatement of the same graph: differences are summation-order only
    assert np.allclose(loc.numpy()[:, ::ROW_STRIDE], g['loc'], rtol=0, atol=2e-5)
    assert np.allclose(conf.numpy()[:, ::ROW_STRIDE], g['conf'], rtol=0, atol=2e-6)
    assert np.allclose(obj.numpy()[:, ::ROW_STRIDE], g['obj'], rtol=0, atol=2e-6)
    assert np.allclose(checksum(conf), g['conf_sum'], rtol=1e-5)


@pytest.mark.skipif(not ref_import.available(), reason='reference not present (GPU box)')
def test_oracle_against_live_reference():
    r = ref_import.load()
    d = synth.random_dets(500, seed=11)
    assert np_oracle.nms(d, 0.3, False) == list(r.py_cpu_nms(d, 0.3))
    if r.cpu_nms is not None:
        assert c_oracle.cpu_nms(d, 0.3, True) == list(r.cpu_nms(d, 0.3))
    priors = r.PriorBox(r.cfg.COCO_300).forward()
    loc = torch.randn(priors.size(0), 4, generator=synth._gen(3, 'loc'))
    want = r.box_utils.decode(loc, priors, [0.1, 0.2]).numpy()
    got = np_oracle.decode(loc.numpy(), priors.numpy())
    assert np.allclose(got, want, rtol=0, atol=2e-6)


def test_init_reweight_prototypes_match_reference(golden):
    """train.py:252-286 restated (np_oracle.init_reweight_prototypes + np_oracle.match) against the golden produced by
    executing the reference's statements with the reference's own match()."""
    from oracle.gen_golden import reweight_inputs
    g = golden('reweight.npz')
    priors = np_oracle.prior_box(ctx.VOC_300)
    P = priors.shape[0]
    feats, targets_all = reweight_inputs(P)
    labels = []
    for targets in targets_all:
        lab = np.zeros((len(targets), P), np.float32)
        for i, t in enumerate(targets):
            t = t.numpy()
            lab[i] = np_oracle.match(0.5, t[:, :4], priors, (0.1, 0.2), t[:, 4:6])[1][:, 0]
        labels.append(lab)
    assert np.array_equal(np.array([checksum(torch.from_numpy(l)) for l in labels]), g['labels_checksum'])
    got = np_oracle.init_reweight_prototypes([f.numpy() for f in feats], labels, 21, 'transfer')
    assert np.array_equal(np.isnan(got), np.isnan(g['weight']))
    assert np.isnan(g['weight']).any(1).sum() == 10 and (g['counts'] > 0).sum() == 10        # both cases are covered
    assert np.allclose(got, g['weight'], rtol=0, atol=2e-6, equal_nan=True)
    inc = np_oracle.init_reweight_prototypes([f.numpy() for f in feats], labels, 21, 'incre')
    assert inc.shape == (5, 60) and np.allclose(inc, g['weight_incre'], rtol=0, atol=2e-6, equal_nan=True)


def test_base_transform_resize_oracle_vs_reference_golden(golden):
    """oracle/np_oracle.base_transform (OpenCV's fixed-point 8-bit bilinear resize restated) against the output of the
    REFERENCE's own BaseTransform class (data/data_augment.py:224-266) on seeded images — tests/golden/resize.npz."""
    from oracle.gen_golden import RESIZE_CASES, resize_image
    g = golden('resize.npz')
    for hw, size, seed in RESIZE_CASES:
        key = '%dx%d_%d' % (hw[0], hw[1], size)
        t = np_oracle.base_transform(resize_image(hw, seed), size)
        assert t.shape == (3, size, size) and t.dtype == np.float32
        assert np.array_equal(t[:, ::17, :], g['rows_' + key]), key
        assert np.array_equal(checksum(t), g['sum_' + key]), key


def test_resize_oracle_vs_live_cv2():
    """... and against cv2.resize itself where cv2 is importable: up / down scaling, odd sizes, 1-pixel sources."""
    cv2 = pytest.importorskip('cv2')
    rng = np.random.default_rng(7)
    for (h, w) in [(375, 500), (500, 375), (333, 500), (300, 300), (600, 600), (100, 130), (1, 1), (2, 3), (480, 640), (299, 301), (37, 901)]:
        for size in (300, 512):
            img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
            assert np.array_equal(np_oracle.cv_resize_linear_u8(img, size), cv2.resize(img, (size, size), interpolation=cv2.INTER_LINEAR)), (h, w, size)

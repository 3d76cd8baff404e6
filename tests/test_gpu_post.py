"""GPU parity: Detect / NMS / soft-NMS / fused post-processing / match / ranking / loss through the
C ABI against the CPU oracle and the golden vectors of the real reference.
Bar: kept indices, prior indices, class ids, labels, ranks bit-exact; floats within the stated
tolerance (decode uses expf, encode uses logf: <= 2 ulp from the CPU libm)."""
import ctypes

import numpy as np
import pytest
import torch

import context_transformer_b200 as ctx
from context_transformer_b200 import _lib
from oracle import c_oracle, np_oracle, synth
from oracle.gen_golden import ROW_STRIDE

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _heads(B, P, C, seed=0):
    return synth.calibrated_heads(B, P, C, seed=seed)


# ---------------------------------------------------------------------------------------------
def test_detect_forward_vs_oracle_and_golden(golden):
    g = golden('post_voc300.npz')
    priors = ctx.PriorBox(ctx.VOC_300).forward()
    loc, conf, obj = _heads(2, priors.size(0), 20)
    det = ctx.Detect(21, 0, ctx.VOC_300)
    boxes, scores = det.forward((loc.to(DEV), conf.to(DEV), obj.to(DEV)), priors.to(DEV))
    assert boxes.is_cuda and tuple(boxes.shape) == (2, 11620, 4) and tuple(scores.shape) == (2, 11620, 21)
    assert det.boxes is boxes and det.scores is scores
    ob, os_ = np_oracle.detect(loc.numpy(), conf.numpy(), obj.numpy(), priors.numpy())
    assert np.array_equal(scores.cpu().numpy(), os_)                       # one fp32 multiply: exact
    assert np.allclose(boxes.cpu().numpy(), ob, rtol=0, atol=2e-6)         # expf vs libm exp
    assert np.allclose(boxes.cpu().numpy()[:, ::ROW_STRIDE], g['boxes'], rtol=0, atol=2e-6)
    assert np.array_equal(scores.cpu().numpy()[:, ::ROW_STRIDE], g['scores'])


def test_decode_helper():
    priors = ctx.PriorBox(ctx.COCO_300).forward()
    loc = torch.randn(priors.size(0), 4, generator=synth._gen(5, 'loc'))
    got = ctx.decode(loc.to(DEV), priors.to(DEV), [0.1, 0.2]).cpu().numpy()
    assert np.allclose(got, np_oracle.decode(loc.numpy(), priors.numpy()), rtol=0, atol=2e-6)


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('n', [1, 2, 63, 64, 65, 300, 2000, 4096, 4097, 9000])
def test_hard_nms_bit_exact(golden, n):
    seed = {0: 0, 1: 1, 63: 2, 64: 3, 65: 4, 300: 5, 2000: 6}.get(n, n)
    d = synth.random_dets(n, seed=seed)
    for on_equal in (False, True):
        got = ctx.nms(d, 0.45, force_cpu=on_equal)
        assert got == c_oracle.cpu_nms(d, 0.45, on_equal)
        key = 'keep_%s_%d' % ('ge' if on_equal else 'gt', n)
        g = golden('nms.npz')
        if key in g.files:
            assert got == g[key].tolist()
    dev_keep = ctx.nms_device(torch.from_numpy(d).to(DEV), 0.45).cpu().tolist()
    assert dev_keep == c_oracle.cpu_nms(d, 0.45, False)


def test_hard_nms_threshold_sweep_and_ties():
    d = synth.random_dets(1500, seed=21)
    for t in (0.0, 0.1, 0.3, 0.7, 1.0):
        for on_equal in (False, True):
            assert ctx.nms(d, t, force_cpu=on_equal) == c_oracle.cpu_nms(d, t, on_equal)
    # ties in score: order is (score desc, index asc) in both the oracle and the kernel
    d2 = synth.random_dets(800, seed=22, tie_free=False)
    d2[:, 4] = np.round(d2[:, 4] * 20) / 20
    assert ctx.nms(d2, 0.45) == c_oracle.cpu_nms(d2, 0.45, False)
    # identical boxes: IoU == 1 exactly; ">= 1.0" suppresses, "> 1.0" does not
    d3 = np.tile(np.array([[10, 10, 50, 60, 0.5]], np.float32), (5, 1))
    d3[:, 4] = [0.5, 0.9, 0.7, 0.6, 0.8]
    assert ctx.nms(d3, 1.0, force_cpu=True) == [1]
    assert ctx.nms(d3, 1.0, force_cpu=False) == [1, 4, 2, 3, 0]


def test_legacy_nms_abi_presorted():
    """void _nms(keep, num, boxes_host, n, dim, thresh, device) — gpu_nms.hpp:1-2 contract."""
    d = synth.random_dets(700, seed=31)
    order = np.argsort(-d[:, 4], kind='stable')
    sorted_d = np.ascontiguousarray(d[order])
    keep = np.empty(700, np.int32)
    num = ctypes.c_int(-1)
    _lib.lib()._nms(keep.ctypes.data, ctypes.addressof(num), sorted_d.ctypes.data, 700, 5, 0.45, 0)
    got = order[keep[:num.value]].tolist()
    assert got == c_oracle.cpu_nms(d, 0.45, False)


@pytest.mark.parametrize('method', [0, 1, 2])
@pytest.mark.parametrize('n', [1, 2, 65, 300, 1500])
def test_soft_nms_bit_exact(golden, method, n):
    seed = {1: 1, 65: 4, 300: 5}.get(n, n)
    d = synth.random_dets(n, seed=seed)
    want, n_want = c_oracle.cpu_soft_nms(d, 0.5, 0.3, 0.001, method)
    got = d.copy()
    keep = ctx.cpu_soft_nms(got, sigma=0.5, Nt=0.3, threshold=0.001, method=method)
    assert keep == list(range(n_want))
    assert np.array_equal(got[:n_want, :4], want[:, :4])
    if method == 2:      # gaussian weight goes through exp(): double exp on both sides, allow 1 ulp
        assert np.allclose(got[:n_want, 4], want[:, 4], rtol=2e-7, atol=0)
    else:
        assert np.array_equal(got[:n_want, 4], want[:, 4])
    key = 'soft_m%d_%d' % (method, n)
    g = golden('nms.npz')
    if key in g.files and method != 2:
        assert np.array_equal(got[:n_want], g[key])


# ---------------------------------------------------------------------------------------------
def _oracle_records(loc, conf, obj, priors, scale, on_equal, thresh=0.01, nms_thresh=0.45, max_per_image=200):
    boxes, scores = np_oracle.detect(loc, conf, obj, priors)
    out = []
    for b in range(loc.shape[0]):
        sc = scale[b] if np.ndim(scale) == 2 else scale
        dets, idx = np_oracle.postprocess_image(boxes[b], scores[b], sc, thresh, nms_thresh, max_per_image,
                                                nms_fn=lambda d, t: c_oracle.cpu_nms(d, t, on_equal))
        out.append(np_oracle.records_from_dets(dets, idx))
    return out


def _check_records(rec, cnt, pidx, want):
    rec, cnt, pidx = rec.cpu().numpy(), cnt.cpu().numpy(), pidx.cpu().numpy()
    for b, (wrec, widx) in enumerate(want):
        n = len(wrec)
        assert cnt[b] == n
        assert np.array_equal(pidx[b, :n], widx.astype(np.int32))           # which prior, in which order: exact
        assert np.array_equal(rec[b, :n, 5], wrec[:, 5])                    # class ids: exact
        assert np.array_equal(rec[b, :n, 4], wrec[:, 4])                    # scores: one fp32 multiply, exact
        assert np.allclose(rec[b, :n, :4], wrec[:, :4], rtol=0, atol=1e-3)  # pixels (<= 500 * 2e-6)
        assert np.all(pidx[b, n:] == -1)


@pytest.mark.parametrize('on_equal', [False, True])
def test_postprocess_calibrated_vs_oracle_and_golden(golden, on_equal):
    g = golden('post_voc300.npz')
    priors = ctx.PriorBox(ctx.VOC_300).forward()
    loc, conf, obj = _heads(2, priors.size(0), 20)
    post = ctx.DetectPost(21, 0, ctx.VOC_300, suppress_on_equal=on_equal)
    rec, cnt, pidx = post.forward((loc.to(DEV), conf.to(DEV), obj.to(DEV)), priors.to(DEV), g['scale'])
    want = _oracle_records(loc.numpy(), conf.numpy(), obj.numpy(), priors.numpy(), g['scale'], on_equal)
    _check_records(rec, cnt, pidx, want)
    conv = 'ge' if on_equal else 'gt'
    for b in range(2):
        n = int(cnt[b])
        assert np.array_equal(pidx[b, :n].cpu().numpy(), g['prior_idx_%s_%d' % (conv, b)])
        assert np.array_equal(rec[b, :n, 4:].cpu().numpy(), g['records_%s_%d' % (conv, b)][:, 4:])
        assert np.allclose(rec[b, :n, :4].cpu().numpy(), g['records_%s_%d' % (conv, b)][:, :4], rtol=0, atol=1e-3)


def test_postprocess_per_image_scale_and_no_topk():
    priors = ctx.PriorBox(ctx.COCO_300).forward()
    B = 3
    loc, conf, obj = _heads(B, priors.size(0), 20, seed=3)
    scale = np.array([[500, 375, 500, 375], [640, 480, 640, 480], [300, 300, 300, 300]], np.float32)
    post = ctx.DetectPost(21, 0, ctx.COCO_300, max_per_image=0, max_out=16384)
    rec, cnt, pidx = post.forward((loc.to(DEV), conf.to(DEV), obj.to(DEV)), priors.to(DEV), scale)
    want = _oracle_records(loc.numpy(), conf.numpy(), obj.numpy(), priors.numpy(), scale, False, max_per_image=0)
    _check_records(rec, cnt, pidx, want)


def test_postprocess_dense_worst_case_and_empty():
    """Random-weight regime: every prior passes the threshold in every class, and the
    opposite extreme where nothing does."""
    g = synth._gen(9, 'dense')
    B, P, C = 2, 1500, 3
    priors = torch.rand(P, 4, generator=g) * 0.5 + 0.1
    loc = torch.randn(B, P, 4, generator=g) * 0.5
    conf = torch.softmax(torch.randn(B, P, C, generator=g), -1)
    obj = torch.softmax(torch.randn(B, P, 2, generator=g), -1)
    cfg = {'variance': [0.1, 0.2]}
    scale = np.array([500, 375, 500, 375], np.float32)
    post = ctx.DetectPost(C + 1, 0, cfg, score_thresh=0.0, max_per_image=200)
    rec, cnt, pidx = post.forward((loc.to(DEV), conf.to(DEV), obj.to(DEV)), priors.to(DEV), scale)
    want = _oracle_records(loc.numpy(), conf.numpy(), obj.numpy(), priors.numpy(), scale, False, thresh=0.0)
    _check_records(rec, cnt, pidx, want)
    post_none = ctx.DetectPost(C + 1, 0, cfg, score_thresh=2.0)
    rec, cnt, pidx = post_none.forward((loc.to(DEV), conf.to(DEV), obj.to(DEV)), priors.to(DEV), scale)
    assert cnt.cpu().tolist() == [0, 0] and float(rec.abs().sum()) == 0.0


@pytest.mark.parametrize('method', [1, 2, 3])
def test_postprocess_soft_nms(method):
    priors = ctx.PriorBox(ctx.VOC_300).forward()
    loc, conf, obj = _heads(1, priors.size(0), 20, seed=4)
    scale = np.array([500, 375, 500, 375], np.float32)
    post = ctx.DetectPost(21, 0, ctx.VOC_300, nms_thresh=0.3, nms_method=method, max_per_image=0, max_out=16384)
    rec, cnt, pidx = post.forward((loc.to(DEV), conf.to(DEV), obj.to(DEV)), priors.to(DEV), scale)
    # soft-NMS decays scores by a function of the IoU, so it is fed the SAME decoded boxes as the kernel
    # (GPU expf vs libm exp differ by <= 2 ulp in w/h, which would leak into the decayed scores)
    gb, gs = ctx.Detect(21, 0, ctx.VOC_300).forward((loc.to(DEV), conf.to(DEV), obj.to(DEV)), priors.to(DEV))
    boxes, scores = gb.cpu().numpy(), gs.cpu().numpy()
    bx = (boxes[0] * scale).astype(np.float32)
    rows = []
    for j in range(1, 21):
        inds = np.where(scores[0][:, j] > np.float32(0.01))[0]
        if len(inds) == 0:
            continue
        c_dets = np.hstack((bx[inds], scores[0][inds, j][:, None])).astype(np.float32)
        out, n = c_oracle.cpu_soft_nms(c_dets, 0.5, 0.3, 0.001, 0 if method == 3 else method)
        rows.append(np.hstack([out, np.full((n, 1), j, np.float32)]))
    want = np.vstack(rows)
    n = int(cnt[0])
    got = rec[0, :n].cpu().numpy()
    assert n == len(want)
    assert np.array_equal(got[:, 5], want[:, 5])
    assert np.array_equal(got[:, :4], want[:, :4])
    assert np.allclose(got[:, 4], want[:, 4], rtol=3e-7 if method == 2 else 0, atol=0)


@pytest.mark.parametrize('method,soft_threshold', [(1, 0.001), (1, 0.2), (2, 0.2), (3, 0.001)])
def test_postprocess_soft_nms_tied_scores(method, soft_threshold):
    """Scores quantised to sixteenths (the first maximum among equal scores is decided by POSITION, which the swap-with-last
    compaction keeps changing) and a soft threshold above some of the initial scores (boxes that never overlap a selected one
    stay in the list below the threshold, cpu_nms.pyx:139-147 only tests a score it has just decayed): the cases in which the
    selection that follows a compaction without a second pass over the scores could go wrong."""
    priors = ctx.PriorBox(ctx.VOC_300).forward()
    P = priors.size(0)
    g = synth._gen(11, 'ties%d' % method)
    loc = torch.randn(1, P, 4, generator=g) * 0.3
    level = torch.randint(1, 9, (1, P, 20), generator=g).float() / 16.0
    conf = torch.where(torch.rand(1, P, 20, generator=g) < 0.08, level, torch.zeros(()))
    conf[..., 3] = torch.where(torch.rand(1, P, generator=g) < 0.4, level[..., 3], torch.zeros(()))     # one long list (generic kernel)
    obj = torch.zeros(1, P, 2)
    obj[..., 1] = 1.0
    scale = np.array([500, 375, 500, 375], np.float32)
    post = ctx.DetectPost(21, 0, ctx.VOC_300, nms_thresh=0.3, nms_method=method, soft_threshold=soft_threshold, max_per_image=0,
                          max_out=65536)
    rec, cnt, pidx = post.forward((loc.to(DEV), conf.to(DEV), obj.to(DEV)), priors.to(DEV), scale)
    gb, gs = ctx.Detect(21, 0, ctx.VOC_300).forward((loc.to(DEV), conf.to(DEV), obj.to(DEV)), priors.to(DEV))
    boxes, scores = gb.cpu().numpy(), gs.cpu().numpy()
    bx = (boxes[0] * scale).astype(np.float32)
    rows, lens = [], []
    for j in range(1, 21):
        inds = np.where(scores[0][:, j] > np.float32(0.01))[0]
        lens.append(len(inds))
        if len(inds) == 0:
            continue
        c_dets = np.hstack((bx[inds], scores[0][inds, j][:, None])).astype(np.float32)
        out, n = c_oracle.cpu_soft_nms(c_dets, 0.5, 0.3, soft_threshold, 0 if method == 3 else method)
        rows.append(np.hstack([out, np.full((n, 1), j, np.float32)]))
    assert min(lens) > 300 and max(lens) > 2048                       # both the shared-memory kernel and the generic one ran
    want = np.vstack(rows)
    n = int(cnt[0])
    got = rec[0, :n].cpu().numpy()
    assert n == len(want)
    assert np.array_equal(got[:, 5], want[:, 5])
    assert np.array_equal(got[:, :4], want[:, :4])
    assert np.allclose(got[:, 4], want[:, 4], rtol=3e-7 if method == 2 else 0, atol=0)


def test_postprocess_full_size_properties():
    """BASELINE config 3 size (512x512 priors, B = 16): properties the oracle need not be run for."""
    priors = ctx.PriorBox(ctx.VOC_512).forward()
    B, P = 16, priors.size(0)
    loc, conf, obj = _heads(B, P, 20, seed=7)
    scale = np.array([512, 512, 512, 512], np.float32)
    post = ctx.DetectPost(21, 0, ctx.VOC_512)
    pred = (loc.to(DEV), conf.to(DEV), obj.to(DEV))
    rec, cnt, pidx = post.forward(pred, priors.to(DEV), scale)
    rec2, cnt2, pidx2 = post.forward(pred, priors.to(DEV), scale)
    assert torch.equal(rec, rec2) and torch.equal(cnt, cnt2) and torch.equal(pidx, pidx2)      # deterministic
    rec, cnt, pidx = rec.cpu().numpy(), cnt.cpu().numpy(), pidx.cpu().numpy()
    boxes, scores = ctx.Detect(21, 0, ctx.VOC_512).forward(pred, priors.to(DEV))
    boxes, scores = boxes.cpu().numpy(), scores.cpu().numpy()
    for b in range(B):
        n = int(cnt[b])
        assert 0 < n <= post.max_out
        r = rec[b, :n]
        cls = r[:, 5].astype(int)
        assert np.all(np.diff(cls) >= 0)                                   # class ascending
        for j in np.unique(cls):
            s = r[cls == j, 4]
            assert np.all(np.diff(s) <= 0)                                 # score descending within a class
            # idempotence: survivors of NMS do not suppress each other
            keep = c_oracle.cpu_nms(np.ascontiguousarray(r[cls == j, :5]), 0.45, False)
            assert keep == list(range(len(s)))
        # every record is the decoded prior it claims to be, with that prior's score
        assert np.allclose(r[:, :4], boxes[b, pidx[b, :n]] * scale, rtol=0, atol=1e-3)
        assert np.array_equal(r[:, 4], scores[b, pidx[b, :n], cls])
        assert np.all(r[:, 4] > np.float32(0.01))
    # one image against the oracle end to end
    want = _oracle_records(loc.numpy()[:1], conf.numpy()[:1], obj.numpy()[:1], priors.numpy(), scale, False)
    _check_records(torch.from_numpy(rec[:1]), torch.from_numpy(cnt[:1]), torch.from_numpy(pidx[:1]), want)


# ---------------------------------------------------------------------------------------------
def test_match_vs_oracle_and_golden(golden):
    g = golden('match_loss.npz')
    priors = ctx.PriorBox(ctx.VOC_300).forward()
    targets = [torch.from_numpy(np.asarray(t, dtype=np.float32)) for t in g['targets']]
    loc_t, conf_t, obj_t, bti, ovl = ctx.match_batch(0.5, targets, priors.to(DEV), [0.1, 0.2], want_overlap=True)
    assert np.array_equal(conf_t.cpu().numpy(), g['conf_t'])
    assert np.array_equal(obj_t.cpu().numpy(), g['obj_t'])
    assert np.allclose(ovl.cpu().numpy()[:, ::ROW_STRIDE], g['overlap'], rtol=0, atol=1e-6)
    pos = conf_t[:, :, 0] != 0
    assert np.array_equal(pos.nonzero().cpu().numpy().astype(np.int32), g['pos_index'])
    assert np.allclose(loc_t[pos].cpu().numpy(), g['loc_t_pos'], rtol=0, atol=2e-5)
    for i, t in enumerate(targets):
        l, c, o, idx, raw = np_oracle.match(0.5, t[:, :4].numpy(), priors.numpy(), (0.1, 0.2), t[:, 4:6].numpy())
        assert np.array_equal(bti[i].cpu().numpy(), idx.astype(np.int32))
        assert np.array_equal(ovl[i].cpu().numpy(), raw)                  # IoU in reference op order: exact
        assert np.allclose(loc_t[i].cpu().numpy(), l, rtol=0, atol=2e-5)  # logf vs libm log
    # reference in-place signature
    B, P = len(targets), priors.size(0)
    lt = torch.zeros(B, P, 4, device=DEV)
    ct = torch.zeros(B, P, 2, device=DEV)
    ot = torch.zeros(B, P, dtype=torch.bool, device=DEV)
    ctx.match(0.5, targets[1][:, :4].to(DEV), priors.to(DEV), [0.1, 0.2], targets[1][:, 4:6].to(DEV), lt, ct, ot, 1)
    assert torch.equal(ct[1], conf_t[1]) and torch.equal(ot[1], obj_t[1]) and torch.equal(lt[1], loc_t[1])
    assert float(ct[0].abs().sum()) == 0.0


def test_match_collisions_and_many_objects():
    """Two ground truths sharing a best prior (last one wins, box_utils.py:122-123), 40 objects."""
    priors = ctx.PriorBox(ctx.VOC_300).forward()
    t = torch.tensor([[0.30, 0.30, 0.60, 0.60, 3, 1.0], [0.301, 0.301, 0.601, 0.601, 7, 0.5]])
    many = synth.synthetic_targets(1, seed=5, max_obj=4)[0].repeat(10, 1)
    many[:, :4] += torch.rand(many.size(0), 4, generator=synth._gen(1, 'jit')) * 0.05
    many[:, 2:4] = torch.maximum(many[:, 2:4], many[:, :2] + 0.05)
    for tg in (t, many):
        loc_t, conf_t, obj_t, bti = ctx.match_batch(0.5, [tg], priors.to(DEV), [0.1, 0.2])
        l, c, o, idx, _ = np_oracle.match(0.5, tg[:, :4].numpy(), priors.numpy(), (0.1, 0.2), tg[:, 4:6].numpy())
        assert np.array_equal(bti[0].cpu().numpy(), idx.astype(np.int32))
        assert np.array_equal(conf_t[0].cpu().numpy(), c)
        assert np.array_equal(obj_t[0].cpu().numpy(), o)


@pytest.mark.parametrize('P', [100, 4096, 11620, 32756])
def test_hard_negative_rank(P):
    g = synth._gen(P, 'rank')
    loss = torch.rand(3, P, generator=g)
    loss[0, ::7] = 0.0                     # ties (positives are zeroed upstream)
    loss[1] = torch.round(loss[1] * 50) / 50
    rank = ctx.hard_negative_rank(loss.to(DEV)).cpu().numpy()
    assert np.array_equal(rank, np_oracle.hard_negative_rank(loss.numpy()).astype(np.int32))


def test_multibox_loss_vs_golden_and_oracle(golden):
    """Fused loss forward + backward (ctx_loss_mining / ctx_hard_negative_rank / ctx_loss_forward_backward) against the
    reference module's own values AND its autograd gradients (tests/golden/match_loss.npz: d(1*box + 2*cls + 3*obj) / d
    loc, conf, obj from /root/reference/layers/modules/multibox_loss_combined.py run on CPU)."""
    g = golden('match_loss.npz')
    priors = ctx.PriorBox(ctx.VOC_300).forward()
    targets = [torch.from_numpy(np.asarray(t, dtype=np.float32)) for t in g['targets']]
    B, P = len(targets), priors.size(0)
    gen = synth._gen(0, 'losspred')
    loc_p = torch.randn(B, P, 4, generator=gen).to(DEV).requires_grad_()
    conf_p = torch.randn(B, P, 20, generator=gen).to(DEV).requires_grad_()
    obj_p = torch.randn(B, P, 2, generator=gen).to(DEV).requires_grad_()
    crit = ctx.MultiBoxLoss_combined(21, 0.5, True, 0, True, 3, 0.5, False)
    out = crit((loc_p, conf_p, obj_p), priors.to(DEV), targets)
    got = np.array([float(out['loss_box_reg']), float(out['loss_cls']), float(out['loss_obj'])])
    assert np.allclose(got, g['loss'], rtol=2e-5)
    wl, wc, wo = (float(v) for v in g['grad_weights'])
    (wl * out['loss_box_reg'] + wc * out['loss_cls'] + wo * out['loss_obj']).backward()
    for name, t in (('grad_loc', loc_p), ('grad_conf', conf_p), ('grad_obj', obj_p)):
        want = g[name]
        gt = t.grad.cpu().numpy()
        assert np.array_equal((gt != 0).any(-1), (want != 0).any(-1)), name    # exactly the reference's pos | neg rows
        assert np.allclose(gt, want, rtol=2e-5, atol=1e-8), (name, np.abs(gt - want).max())
    # targets already on the device (no host round trip), same result
    out2 = crit((loc_p.detach(), conf_p.detach(), obj_p.detach()), priors.to(DEV), [t.to(DEV) for t in targets])
    assert all(float(out2[k]) == float(out[k]) for k in out)


def test_box_algebra_helpers_vs_oracle():
    """point_form / jaccard / encode as stand-alone device ops against the numpy oracle (utils/box_utils.py:5-68, 135-156)."""
    from oracle import np_oracle
    g = synth._gen(5, 'boxalg')
    pri = ctx.PriorBox(ctx.VOC_300).forward()[::7].contiguous()
    xy = torch.rand(9, 2, generator=g) * 0.5
    tr = torch.cat([xy, xy + 0.1 + 0.4 * torch.rand(9, 2, generator=g)], 1)
    pf = ctx.point_form(pri.to(DEV)).cpu().numpy()
    assert np.array_equal(pf, np_oracle.point_form(pri.numpy()))
    assert np.allclose(ctx.jaccard(tr.to(DEV), torch.from_numpy(pf).to(DEV)).cpu().numpy(), np_oracle.jaccard(tr.numpy(), pf), rtol=0, atol=1e-7)
    matched = tr[torch.randint(0, 9, (pri.size(0),), generator=g)]
    assert np.allclose(ctx.encode(matched.to(DEV), pri.to(DEV), [0.1, 0.2]).cpu().numpy(), np_oracle.encode(matched.numpy(), pri.numpy(), (0.1, 0.2)),
                       rtol=2e-5, atol=1e-6)


def test_init_reweight_vs_golden_and_oracle(golden):
    """OBJ(Target) prototype initialisation (train.py:252-286): ctx_match_encode + ctx_prototype_accumulate / _finalize through
    the reference-shaped ``init_reweight(args, model, data_loader, ...)`` against the golden of the reference's own statements.
    Tolerance 2e-6 absolute on unit-norm rows (fp64 per-class sums here, fp32 ``mean`` upstream); NaN rows for classes
    without a sample, exactly where upstream has them."""
    import types
    from oracle.gen_golden import reweight_inputs
    g = golden('reweight.npz')
    priors = ctx.PriorBox(ctx.VOC_300).forward()
    feats, targets_all = reweight_inputs(priors.size(0))

    class FakeNet(torch.nn.Module):
        """Stands in for RFBNet: ``model(data, init=True)`` returns the seeded raw conf features of that batch."""
        def __init__(self):
            super().__init__()
            self.OBJ_Target = torch.nn.Linear(60, 20, bias=False)
            self.calls = 0

        def forward(self, x, init=False):
            assert init
            self.calls += 1
            return feats[self.calls - 1].to(x.device)

    net = FakeNet().cuda()
    loader = [(torch.zeros(4, 3, 8, 8), t) for t in targets_all] + [(torch.zeros(4, 3, 8, 8), targets_all[0])] * 2
    w = ctx.init_reweight(types.SimpleNamespace(init_iter=3, setting='transfer'), net, loader, priors, 21, 0.5)
    torch.cuda.synchronize()
    assert net.calls == 3 and tuple(w.shape) == (20, 60) and net.OBJ_Target.weight.data_ptr() == w.data_ptr()
    got = w.cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(g['weight']))
    assert np.allclose(got, g['weight'], rtol=0, atol=2e-6, equal_nan=True)
    # 'incre': only the last five classes become OBJ(Target) rows (train.py:281-282)
    net2 = FakeNet().cuda()
    w2 = ctx.init_reweight(types.SimpleNamespace(init_iter=50, setting='incre'), net2, loader[:3], priors, 21, 0.5)
    assert tuple(w2.shape) == (5, 60) and np.allclose(w2.cpu().numpy(), g['weight_incre'], rtol=0, atol=2e-6, equal_nan=True)
    # the accumulator alone: counts per class are exact
    acc = ctx.PrototypeAccumulator(20, 60, 'cuda:0')
    for f, t in zip(feats, targets_all):
        acc.add(f.cuda(), ctx.match_batch(0.5, t, priors.cuda(), (0.1, 0.2))[1])
    assert np.array_equal(acc.counts.cpu().numpy(), g['counts'])
    with pytest.raises(_lib.CtxError):
        ctx.PrototypeAccumulator(20, 60, 'cpu')


def test_base_transform_with_resize_bit_exact(golden):
    """On-device BaseTransform incl. the cv2-style 8-bit bilinear resize (ctx_base_transform_resize) — bit-exact against the
    numpy oracle and the golden of the reference's own BaseTransform class, for the mixed image sizes a VOC loop sees."""
    from oracle.gen_golden import RESIZE_CASES, resize_image
    g = golden('resize.npz')
    for hw, size, seed in RESIZE_CASES:
        img = resize_image(hw, seed)
        tr = ctx.BaseTransform(size, (104, 117, 123))
        got = tr(torch.from_numpy(img)).cpu().numpy()
        key = '%dx%d_%d' % (hw[0], hw[1], size)
        assert np.array_equal(got, np_oracle.base_transform(img, size)), key
        assert np.array_equal(got[:, ::17, :], g['rows_' + key]), key
    # a list of images of different sizes -> one batch for the network
    imgs = [torch.from_numpy(resize_image(hw, seed)) for hw, size, seed in RESIZE_CASES[:3]]
    batch = ctx.BaseTransform(300, (104, 117, 123)).batch(imgs).cpu().numpy()
    for i, (hw, size, seed) in enumerate(RESIZE_CASES[:3]):
        assert np.array_equal(batch[i], np_oracle.base_transform(resize_image(hw, seed), 300))

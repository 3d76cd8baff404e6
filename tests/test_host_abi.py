"""CPU: host-side logic of the product package and the C-ABI library (load + exports only)."""
import ctypes
import os
import re
import types

import numpy as np
import pytest
import torch

import context_transformer_b200 as ctx
from context_transformer_b200 import _lib, engine, shard
from oracle import synth
from oracle.gen_golden import ROW_STRIDE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, 'include', 'ctx_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    names = re.findall(r'^\s*(?:[\w\s\*]+?)\b(\w+)\s*\([^;{]*\)\s*;', text, flags=re.M)
    return sorted(set(n for n in names if n.startswith('ctx_') or n == '_nms'))


def test_library_exports_every_declared_symbol():
    _lib_path = _lib.LIB_PATH
    if not os.path.exists(_lib_path):
        from context_transformer_b200 import build
        build.build_library()
    declared = _header_functions()
    assert len(declared) >= 35
    raw = ctypes.CDLL(_lib_path)
    for name in declared:
        assert hasattr(raw, name), 'libctx_b200.so does not export %s' % name
    assert sorted(_lib.SIGNATURES) == declared
    L = _lib.lib()
    assert L.ctx_version() >= 100
    assert L.ctx_last_error() is not None
    # pure-host size queries (no GPU needed)
    assert L.ctx_postprocess_workspace_bytes(2, 11620, 20) > 2 * 11620 * 16
    assert L.ctx_nms_workspace_bytes(1000) >= 1024 * 8
    assert L.ctx_rank_workspace_bytes(4, 11620) >= 4 * 16384 * 8


def test_struct_layouts_match_header_sizes(tmp_path):
    """sizeof / last-field offset of every struct as the C compiler sees include/ctx_b200.h == the ctypes mirrors."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = ['CtxPostParams', 'CtxOutSeg', 'CtxConvParams', 'CtxPoolParams', 'CtxAttnParams']
    last = {n: getattr(_lib, n)._fields_[-1][0] for n in names}
    src = tmp_path / 'sz.c'
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ctx_b200.h"\nint main(void) {\n' +
                   ''.join('  printf("%%zu %%zu\\n", sizeof(%s), offsetof(%s, %s));\n' % (n, n, last[n]) for n in names) + '  return 0;\n}\n')
    exe = tmp_path / 'sz'
    subprocess.check_call(['gcc', '-I', os.path.join(root, 'include'), str(src), '-o', str(exe)])
    lines = subprocess.check_output([str(exe)]).decode().split('\n')
    for n, line in zip(names, lines):
        size, off = (int(v) for v in line.split())
        cls = getattr(_lib, n)
        assert ctypes.sizeof(cls) == size, n
        assert getattr(cls, last[n]).offset == off, n


@pytest.mark.parametrize('name', ['VOC_300', 'VOC_512', 'COCO_300', 'COCO_512'])
def test_prior_box_bit_exact(golden, name):
    cfg = getattr(ctx, name)
    p = ctx.PriorBox(cfg).forward()
    assert p.dtype == torch.float32 and tuple(p.shape) == (ctx.num_priors(cfg), 4)
    assert np.array_equal(p.numpy(), golden('priors.npz')[name])


def test_prior_box_rejects_bad_variance():
    cfg = dict(ctx.VOC_300)
    cfg['variance'] = [0.1, -0.2]
    with pytest.raises(ValueError):
        ctx.PriorBox(cfg)


def test_build_net_surface():
    assert ctx.build_net(types.SimpleNamespace(method='ours', phase=2, setting='transfer'), 400, 60) is None
    net = ctx.build_net(types.SimpleNamespace(method='ours', phase=2, setting='incre'), 300, 15)
    assert net.size == 300 and net.indicator == 3
    for name in ('base', 'Norm', 'extras', 'loc', 'conf', 'obj', 'theta', 'phi', 'g', 'OBJ_Target', 'fc_base'):
        assert hasattr(net, name)
    assert tuple(net.Wz.shape) == (15,) and float(net.scale) == 5.0 and not net.scale.requires_grad
    net.OBJ_Target.weight.data.normal_()
    net.normalize()
    assert torch.allclose(net.OBJ_Target.weight.norm(dim=1), torch.ones(5), atol=1e-6)
    # LR groups of utils/solver.py key on these substrings
    names = [n for n, _ in net.named_parameters()]
    assert any(n.startswith('base.') for n in names) and any(n.startswith('extras.') for n in names)
    assert any(n.startswith('Norm.') for n in names)


def test_eval_forward_refuses_cpu():
    net = ctx.build_net(types.SimpleNamespace(method='ft', phase=2, setting='transfer'), 300, 20).eval()
    x = torch.zeros(1, 3, 300, 300)
    with pytest.raises(AttributeError):
        net(x)                                    # device not assigned by the caller yet
    net.device = 'cpu'
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        net(x)


def test_autograd_forward_init_matches_reference(golden):
    """init=True prototype-extraction path (train.py:252-286) — autograd expression, runs anywhere."""
    g = golden('net_ours_transfer_300.npz')
    net = ctx.build_net(types.SimpleNamespace(method='ours', phase=2, setting='transfer'), 300, 60)
    net.load_state_dict(synth.seeded_state(net.state_dict(), seed=0))
    net.eval()
    net.device = 'cpu'
    with torch.no_grad():
        conf = net(synth.seeded_input(2, 300, seed=0), init=True)
    assert np.allclose(conf.numpy()[:, ::ROW_STRIDE], g['conf_init'], rtol=0, atol=2e-5)


def test_ours_at_512_raises_like_reference():
    net = ctx.build_net(types.SimpleNamespace(method='ours', phase=2, setting='transfer'), 512, 60)
    net.device = 'cpu'
    net.train()
    with pytest.raises(IndexError):
        net(torch.zeros(2, 3, 512, 512))


def test_pool_out_ceil_mode():
    for h, k, s, pad, ceil_mode in [(75, 2, 2, 0, True), (38, 3, 3, 0, True), (19, 2, 2, 0, True), (10, 2, 2, 0, True),
                                    (5, 2, 2, 0, True), (3, 1, 1, 0, True), (19, 3, 1, 1, False), (300, 2, 2, 0, False)]:
        want = torch.nn.functional.max_pool2d(torch.zeros(1, 1, h, h), k, s, pad, ceil_mode=ceil_mode).shape[-1]
        assert engine._pool_out(h, k, s, pad, ceil_mode) == want


def test_nms_wrapper_empty_and_validation():
    assert ctx.nms(np.zeros((0, 5), np.float32), 0.45) == []
    assert ctx.nms(np.zeros((0, 5), np.float32), 0.45, force_cpu=True) == []
    with pytest.raises(ValueError):
        ctx.cpu_soft_nms(np.zeros((3, 4), np.float32))
    if not torch.cuda.is_available():
        with pytest.raises(_lib.CtxError):
            ctx.nms(np.ones((3, 5), np.float32), 0.45)


def test_detect_requires_cuda_tensors():
    if torch.cuda.is_available():
        pytest.skip('CPU-only check')
    det = ctx.Detect(21, 0, ctx.VOC_300)
    with pytest.raises(_lib.CtxError):
        det.forward((torch.zeros(1, 8, 4), torch.zeros(1, 8, 20), torch.zeros(1, 8, 2)), torch.zeros(8, 4))


def test_shard_bounds_and_packing():
    assert shard.shard_bounds(256, 8, 3) == (96, 128)
    with pytest.raises(ValueError):
        shard.shard_bounds(10, 4, 0)
    rec = torch.randn(3, 7, 6)
    cnt = torch.tensor([0, 7, 11620 * 20], dtype=torch.int32)
    r2, c2 = shard.unpack_records(shard.pack_records(rec, cnt))
    assert torch.equal(r2, rec) and torch.equal(c2, cnt)
    r3, c3 = shard.gather_records(rec, cnt)          # no process group: identity
    assert torch.equal(r3, rec) and torch.equal(c3, cnt)


def test_records_to_all_boxes():
    rec = torch.zeros(1, 5, 6)
    rec[0, 0] = torch.tensor([1., 2, 3, 4, .9, 3])
    rec[0, 1] = torch.tensor([5., 6, 7, 8, .8, 3])
    rec[0, 2] = torch.tensor([9., 9, 9, 9, .7, 17])
    ab = ctx.records_to_all_boxes(rec, torch.tensor([3]), 21)
    assert ab[3][0].shape == (2, 5) and ab[17][0].shape == (1, 5) and ab[1][0].shape == (0, 5)
    assert ab[3][0][1, 4] == np.float32(.8)


def test_detection_collector_builds_the_reference_result_structure(tmp_path):
    """test.py:107-108,150-154,171-172: all_boxes[class][image] float32 [k,5] arrays (class 0 untouched), pickled."""
    import pickle
    import numpy as np
    import torch
    import context_transformer_b200 as ctx
    K = 8
    rec = torch.zeros(3, K, 6)
    cnt = torch.tensor([3, 0, 11], dtype=torch.int32)                 # image 2 kept more than max_out rows: truncated to K
    rows0 = torch.tensor([[1, 2, 3, 4, .9, 2], [5, 6, 7, 8, .5, 2], [0, 0, 9, 9, .7, 7]])
    rec[0, :3] = rows0
    rec[2] = torch.arange(K * 6, dtype=torch.float32).view(K, 6)
    rec[2, :, 5] = torch.tensor([1, 1, 3, 3, 3, 20, 20, 20])
    col = ctx.DetectionCollector(num_images=5, num_classes=21)
    col.add(1, rec, cnt)
    ab = col.all_boxes
    assert all(ab[0][i] == [] for i in range(5))                       # background class keeps the reference's empty lists
    assert ab[2][1].dtype == np.float32 and np.array_equal(ab[2][1], rows0[:2, :5].numpy())
    assert np.array_equal(ab[7][1], rows0[2:3, :5].numpy())
    assert all(ab[j][2].shape == (0, 5) for j in range(1, 21))         # image with no detection: empty [0,5] arrays
    assert [ab[j][3].shape[0] for j in (1, 3, 20)] == [2, 3, 3] and np.array_equal(ab[3][3], rec[2, 2:5, :5].numpy())
    assert ab[5][0] == [] and ab[5][4] == []                           # images not yet processed
    f = col.save(str(tmp_path / 'detections.pkl'))
    back = pickle.load(open(f, 'rb'))
    assert np.array_equal(back[20][3], ab[20][3]) and back[0][0] == []
    import pytest
    with pytest.raises(IndexError):
        col.add(4, rec, cnt)


def test_result_writers_match_the_reference_writers(golden, tmp_path):
    """f-2: the VOC per-class text files and the COCO result list, written from the gathered [B, K, 6] records through
    DetectionCollector, against tests/golden/writers.npz — produced by EXECUTING the reference's own writer functions
    (data/voc0712.py:360-376, data/coco.py:242-259; oracle/gen_golden.py gen_writers) on the same detections."""
    import json
    from oracle.gen_golden import writer_inputs
    g, post = golden('writers.npz'), golden('post_voc300.npz')
    _, ids = writer_inputs(post)
    K = 256
    rec = torch.zeros(2, K, 6)
    cnt = torch.zeros(2, dtype=torch.int32)
    for b in range(2):
        r = torch.from_numpy(post['records_gt_%d' % b])
        rec[b, :len(r)] = r
        cnt[b] = len(r)
    col = ctx.DetectionCollector(num_images=2, num_classes=21).add(0, rec, cnt)
    voc_classes = ('__background__', 'aeroplane', 'bicycle', 'bird', 'boat', 'bottle', 'bus', 'car', 'cat', 'chair', 'cow', 'diningtable',
                   'dog', 'horse', 'motorbike', 'person', 'pottedplant', 'sheep', 'sofa', 'train', 'tvmonitor')
    files = col.write_voc_results(ids, voc_classes, str(tmp_path / 'Main'))
    assert len(files) == 20
    n_lines = 0
    for f in files:
        text = open(f).read()
        assert text == str(g['voc_' + f.split('/')[-1]]), f
        n_lines += text.count('\n')
    assert n_lines == int(cnt.sum())
    names = voc_classes[1:]
    res = col.coco_results([139, 285], names, {n: 101 + i for i, n in enumerate(names)})
    assert [x['image_id'] for x in res] == g['coco_image_id'].tolist() and [x['category_id'] for x in res] == g['coco_category_id'].tolist()
    assert np.array_equal(np.array([x['bbox'] for x in res]), g['coco_bbox']) and np.array_equal(np.array([x['score'] for x in res]), g['coco_score'])
    out = col.write_coco_results(str(tmp_path / 'res.json'), [139, 285], names, {n: 101 + i for i, n in enumerate(names)})
    assert len(json.load(open(out))) == len(res)

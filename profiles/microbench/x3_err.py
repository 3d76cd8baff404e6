import os, sys, types
sys.path.insert(0, '' + os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..') + ''); sys.path.insert(0, '' + os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..') + '/tests')
import numpy as np, torch
import context_transformer_b200 as ctx
from oracle import synth
from oracle.gen_golden import NET_CASES, ROW_STRIDE
from test_gpu_net import _build
for prec, attn in (('fp32', ''), ('fp32x3', ''), ('fp32x3', '0'), ('fp32', '2')):
    if attn: os.environ['CTX_ATTN_MODE'] = attn
    else: os.environ.pop('CTX_ATTN_MODE', None)
    for case in NET_CASES:
        tag, method, phase, setting, size, ncls, batch = case
        g = np.load('' + os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..') + '/tests/golden/net_%s.npz' % tag)
        net = _build(case, prec)
        loc, conf, obj = [t.cpu().numpy() for t in net(synth.seeded_input(batch, size, seed=0))]
        dl, dc, do = [np.abs(a[:, ::ROW_STRIDE] - g[k]) for a, k in ((loc, 'loc'), (conf, 'conf'), (obj, 'obj'))]
        rl = (dl / np.maximum(1.0, np.abs(g['loc']))).max()
        print('%-7s attn=%-2s %-18s loc %.2e (rel-to-max(1,|ref|) %.2e, max|loc| %.0f) conf %.2e obj %.2e' % (prec, attn or '-', tag, dl.max(), rl, np.abs(g['loc']).max(), dc.max(), do.max()), flush=True)

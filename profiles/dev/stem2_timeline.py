"""Development aid: clock64 timeline of CTA 0 of the fused conv1_1 + conv1_2 kernel (300x300, batch 32)."""
import sys
import ctypes as C

import torch

sys.path.insert(0, '.')
sys.path.insert(0, 'profiles/dev')
from stem2_check import Scratch, DEV  # noqa
from context_transformer_b200 import _lib
from context_transformer_b200.engine import View


def main():
    N, H, W = 32, 300, 300
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, 3, H, W, generator=g) * 50
    w1 = torch.randn(64, 3, 3, 3, generator=g) * 0.02
    b1 = torch.randn(64, generator=g) * 0.1
    w2 = torch.randn(64, 64, 3, 3, generator=g) * 0.06
    b2 = torch.randn(64, generator=g) * 0.1
    e = Scratch('bf16')
    xr = View(x.to(DEV).contiguous().view(-1), N, H, W, 3)
    mid = View(torch.empty(0, dtype=e.act_dtype, device=DEV), N, H, W, 64)
    out = e._emit_conv('fused', mid, w2.to(DEV), b2.to(DEV), 1, (1, 1), 1, True, pool2=True, stem=(xr, w1.to(DEV), b1.to(DEV)))
    assert out is not None
    for _ in range(3):
        e.run_range(0, 1)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); e.run_range(0, 1, 10); b.record(); b.synchronize()
    print('kernel %.1f us' % (a.elapsed_time(b) * 100))
    buf = torch.zeros(8 * 64 * 6, dtype=torch.int64, device=DEV)
    _lib.lib().ctx_debug_set_conv_timeline(C.c_void_p(buf.data_ptr()))
    e.run_range(0, 1)
    torch.cuda.synchronize()
    _lib.lib().ctx_debug_set_conv_timeline(None)
    t = buf.cpu().view(8, 64, 6)
    t0 = int(t[t > 0].min())
    names = ['tma', 'build', 'issue0', 'issue1', 'mid', 'epi0', 'epi1']
    keys = {'tma': ['issue'], 'build': ['raw_full', 's_full(A1 free)', 'a1_full'], 'issue0': ['acce', 'a2_full', 'main issued', 'a1_full', 'stem issued'],
            'issue1': ['acce', 'a2_full', 'main issued', 'a1_full', 'stem issued'], 'mid': ['accf(A2 free)', 's_full', 'a2_full', 'ld done', 'stores done'], 'epi0': ['accf', 'done', 'ld0', 'chunk0', 'ld1', 'chunk1'], 'epi1': ['accf', 'done', 'ld0', 'chunk0', 'ld1', 'chunk1']}
    for j in list(range(0, 4)) + list(range(20, 26)):
        print('tile %d' % j)
        for ri, nm in enumerate(names):
            vals = [int(v) - t0 for v in t[ri, j] if v > 0]
            if vals:
                print('   %-7s %s' % (nm, '  '.join('%s %d' % (k, v) for k, v in zip(keys[nm], vals))))
    # steady-state period
    m = t[2, :, 2]
    idx = [j for j in range(64) if m[j] > 0]
    if len(idx) > 4:
        print('issuer 0: main-issued period over tiles %d..%d: %.0f clk / tile pair' % (idx[2], idx[-1], float(m[idx[-1]] - m[idx[2]]) / (len(idx) - 3)))


if __name__ == '__main__':
    main()

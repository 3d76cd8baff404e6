"""Development aid: clock64 timeline of CTA 0 of conv_halo_kernel (resident weights) on conv2_1 (64 -> 128 @ 150x150, batch 32)."""
import sys
import ctypes as C

import torch

sys.path.insert(0, '.')
sys.path.insert(0, 'profiles/dev')
from stem2_check import Scratch, DEV  # noqa
from context_transformer_b200 import _lib
from context_transformer_b200.engine import View


def main(cin=64, cout=128, H=150, amode=4):
    N = 32
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, cin, H, H, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) * 0.05
    b = torch.randn(cout, generator=g) * 0.1
    e = Scratch('bf16')
    buf = x.permute(0, 2, 3, 1).contiguous().to(DEV, torch.bfloat16)
    src = View(buf.view(-1), N, H, H, cin)
    out = e._emit_conv('t', src, w.to(DEV), b.to(DEV), 1, (1, 1), 1, True)
    L = e.L
    p = e.last_conv_params
    plan = C.c_void_p()
    _lib.check(L.ctx_conv2d_tc_plan_create_tuned(C.byref(p), 0, 1, amode, 0, C.byref(plan)))
    info = (C.c_int * 8)()
    L.ctx_conv2d_tc_plan_info(plan, info)
    print('plan', list(info))
    st = _lib.current_stream_ptr(DEV)
    for _ in range(3):
        L.ctx_conv2d_tc_plan_run(plan, st)
    torch.cuda.synchronize()
    a, bb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        L.ctx_conv2d_tc_plan_run(plan, st)
    bb.record(); bb.synchronize()
    print('kernel %.1f us' % (a.elapsed_time(bb) * 100))
    tl = torch.zeros(8 * 64 * 6, dtype=torch.int64, device=DEV)
    L.ctx_debug_set_conv_timeline(C.c_void_p(tl.data_ptr()))
    L.ctx_conv2d_tc_plan_run(plan, st)
    torch.cuda.synchronize()
    L.ctx_debug_set_conv_timeline(None)
    t = tl.cpu().view(8, 64, 6)
    t0 = int(t[t > 0].min())
    for j in range(8, 14):
        row = lambda r, ks: ' '.join('%d' % (int(t[r, j if r != 5 and r != 6 else j, k]) - t0) if t[r, j, k] > 0 else '-' for k in ks)
        print('tile %2d | patch issued %s | mma: acce,fullA,issued %s | epi%d: accf,done,ld0,chunk0,ld1,chunk1 %s'
              % (j, row(0, (0,)), row(2, (0, 1, 2)), j & 1, row(5 + (j & 1), (0, 1, 2, 3, 4, 5))))


if __name__ == '__main__':
    main()
    main(128, 128, 150, 3)

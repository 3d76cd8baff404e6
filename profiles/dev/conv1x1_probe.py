"""Development aid: a 1x1 convolution alone (default: the RFB ConvLinear of the 38x38 level, 512 -> 512, batch 32), with and without
the residual input, with and without ReLU, under several tilings: what bounds the short-K layers?"""
import sys
import ctypes as C

import torch

sys.path.insert(0, '.')
sys.path.insert(0, 'profiles/dev')
from stem2_check import Scratch, DEV  # noqa
from context_transformer_b200 import _lib
from context_transformer_b200.engine import View


def run(cin, cout, H, res, tilings, N=32):
    g = torch.Generator().manual_seed(1)
    e = Scratch('bf16')
    buf = (torch.randn(N, H, H, cin, generator=g)).to(DEV, torch.bfloat16)
    src = View(buf.view(-1), N, H, H, cin)
    w = torch.randn(cout, cin, 1, 1, generator=g) * 0.05
    b = torch.randn(cout, generator=g) * 0.1
    rv = None
    if res:
        rbuf = torch.randn(N, H, H, cout, generator=g).to(DEV, torch.bfloat16)
        rv = View(rbuf.view(-1), N, H, H, cout)
    e._emit_conv('t', src, w.to(DEV), b.to(DEV), 1, (0, 0), 1, True, residual=rv)
    L, p = e.L, e.last_conv_params
    st = _lib.current_stream_ptr(DEV)
    for (n, cl, amode, cg) in tilings:
        plan = C.c_void_p()
        if L.ctx_conv2d_tc_plan_create_tuned(C.byref(p), n, cl, amode, cg, C.byref(plan)) != 0:
            continue
        info = (C.c_int * 8)()
        L.ctx_conv2d_tc_plan_info(plan, info)
        for _ in range(3):
            L.ctx_conv2d_tc_plan_run(plan, st)
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            a, bb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20):
                L.ctx_conv2d_tc_plan_run(plan, st)
            bb.record(); bb.synchronize()
            ts.append(a.elapsed_time(bb) * 50)
        gf = 2.0 * N * H * H * cin * cout / 1e9
        print('%dx%d %d->%d res %d  plan %s  %.1f us  %.0f TF/s' % (H, H, cin, cout, int(res), list(info), min(ts), gf / min(ts) * 1e3))
        L.ctx_conv2d_tc_plan_destroy(plan)


if __name__ == '__main__':
    til = [(0, 1, -1, 0), (2, 2, 1, 0), (4, 1, 1, 0), (1, 1, 1, 0)]
    run(512, 512, 38, False, til)
    run(512, 512, 38, True, til)
    run(1024, 1024, 19, False, til)
    run(512, 960, 38, False, til)

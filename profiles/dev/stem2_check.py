"""Development aid: the fused conv1_1 + conv1_2 (+ pool) kernel alone, against torch (16-bit operands, fp32 accumulate)."""
import sys
import ctypes as C

import torch
import torch.nn.functional as F

sys.path.insert(0, '.')
import context_transformer_b200 as ctx  # noqa
from context_transformer_b200 import _lib
from context_transformer_b200.engine import Engine, View

DEV = torch.device('cuda:0')


class Scratch(Engine):
    def __init__(self, precision):
        self.L = _lib.lib()
        self.dev = DEV
        self.precision = precision
        self.act_dtype = {'bf16': torch.bfloat16, 'fp16': torch.float16}[precision]
        self.split = False
        self.act_code = _lib.dtype_code(self.act_dtype)
        self.keep, self.layers = [], []
        self.prog = C.c_void_p()
        _lib.check(self.L.ctx_prog_create(C.byref(self.prog)))


def main(N=2, H=40, W=44, pool=True, precision='bf16'):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, 3, H, W, generator=g) * 50
    w1 = torch.randn(64, 3, 3, 3, generator=g) * 0.02
    b1 = torch.randn(64, generator=g) * 0.1
    w2 = torch.randn(64, 64, 3, 3, generator=g) * 0.06
    b2 = torch.randn(64, generator=g) * 0.1
    e = Scratch(precision)
    dt = e.act_dtype
    xr = View(x.to(DEV).contiguous().view(-1), N, H, W, 3)
    mid = View(torch.empty(0, dtype=dt, device=DEV), N, H, W, 64)
    out = e._emit_conv('fused', mid, w2.to(DEV), b2.to(DEV), 1, (1, 1), 1, True, pool2=pool, stem=(xr, w1.to(DEV), b1.to(DEV)))
    assert out is not None
    e.run_range(0, 1)
    torch.cuda.synchronize()
    got = out.tensor().float().cpu().permute(0, 3, 1, 2)
    a1 = F.relu(F.conv2d(x.to(dt).float(), w1.to(dt).float(), b1, 1, 1)).to(dt).float()
    y = F.relu(F.conv2d(a1, w2.to(dt).float(), b2, 1, 1))
    if pool:
        y = F.max_pool2d(y, 2, 2)
    err = (got - y).abs().max().item()
    print('N %d H %d W %d pool %d %s: max |err| %.4g (scale %.3g)' % (N, H, W, pool, precision, err, y.abs().max().item()))
    assert err < 0.02 * y.abs().max().item()


if __name__ == '__main__':
    main()
    main(3, 300, 300, True, 'bf16')
    main(1, 64, 48, False, 'fp16')

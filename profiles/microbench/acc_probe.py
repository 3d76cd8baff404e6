"""How exact is the fp32 accumulation of tcgen05.mma kind::f16?  Operands exactly representable in 16 bits (products exact
in fp32), fp32 segment output, long K chains: any error against fp64 is accumulation rounding.  Prints the signed mean of
the relative error (a truncating adder shows up as a shrink towards zero) and its rms for several K."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..', 'tests'))
from test_gpu_net import _Scratch, _nhwc_view, DEV

torch.manual_seed(0)
for prec, dt in (('bf16', torch.bfloat16), ('fp16', torch.float16)):
    for cin, k, positive in ((64, 1, 1), (512, 1, 1), (512, 3, 1), (1024, 3, 1), (512, 3, 0)):
        N, H, cout = 2, 19, 128
        x = torch.randn(N, cin, H, H)
        if positive:
            x = x.abs()                                   # post-ReLU like; positive weights too -> monotone partial sums (worst case)
        w = torch.randn(cout, cin, k, k)
        if positive:
            w = w.abs()
        w = w * (1.0 / (cin * k * k))
        x, w = x.to(dt).float(), w.to(dt).float()
        b = torch.zeros(cout)
        e = _Scratch(prec)
        src, _ = _nhwc_view(x, e.act_dtype)
        out = torch.zeros(N, H, H, cout, dtype=torch.float32, device=DEV)
        segs = [(out.view(-1), 0, cout, H * H * cout, cout, 0)]
        e._emit_conv('t', src, w.to(DEV), b.to(DEV), 1, (k // 2, k // 2), 1, False, segs=segs)
        assert e.layers[-1][1] == 'conv_tc'
        e.go()
        got = out.cpu().permute(0, 3, 1, 2).double()
        want = F.conv2d(x.double(), w.double(), None, 1, k // 2)
        w32 = F.conv2d(x, w, None, 1, k // 2).double()
        rel = (got - want) / want.abs().clamp_min(1e-3)
        rel32 = (w32 - want) / want.abs().clamp_min(1e-3)
        print('%s K=%5d positive=%d: tcgen05 mean rel %+.3e rms %.3e max %.3e | cpu fp32 mean %+.3e rms %.3e' % (
            prec, cin * k * k, positive, rel.mean(), rel.pow(2).mean().sqrt(), rel.abs().max(), rel32.mean(), rel32.pow(2).mean().sqrt()))

"""``Engine`` — compiles an ``RFBNet`` (eval mode) into a flat program of sm_100a kernels.

One engine = one (batch, precision, device).  Building it
  * folds every BatchNorm (running stats, eps 1e-5) and bias into a per-channel fp32 epilogue
    vector and packs the conv weights in the kernels' layouts (device-side tensor ops, once);
  * allocates every activation once, NHWC (channels-last), in the engine's precision, so that the
    reference's ``permute(0,2,3,1).contiguous()`` + ``cat`` of the head outputs
    (models/RFB_Net_vgg.py:239-248) is just where the head conv's epilogue writes;
  * turns each RFB block (:26-112) into branch convs whose last conv writes its slice of the
    concat buffer, a shortcut conv, and a ConvLinear conv whose epilogue adds the shortcut and
    applies the ReLU (``out*scale + short`` with ``scale`` folded into the weights);
  * evaluates loc / conf / obj of a level as ONE 3x3 conv with three output segments (the
    reference runs the conf conv twice in phase-2 'ours', :240,243);
  * adds the conf max-pool (ceil_mode, :242-244), the fused Context-Transformer kernel (:253-271)
    and the output softmaxes (:279-285);
and records all of it in a ``ctx_prog`` (csrc/prog.cu), optionally captured into a CUDA graph.
``run(x)`` copies the input into the program's static input buffer and replays the program.
"""
import ctypes as C
import math
import os

import torch
import torch.nn as nn

from . import _lib
from .rfb_net import CONF_POOL, SOURCE_SPLIT, BasicConv, _RFBBlock

_DT = {'fp32': torch.float32, 'bf16': torch.bfloat16, 'fp16': torch.float16, 'fp32x3': torch.float16}
_TC16 = ('bf16', 'fp16')             # plain 16-bit tensor-core modes ('fp32x3' keeps fp16 hi / lo plane PAIRS: fp32 emulated, csrc/conv_x3.cu)


class View(object):
    """A [N,H,W,C] channels-last activation living at channel offset ``coff`` of a buffer whose
    pixels are ``cstride`` channels wide."""
    __slots__ = ('buf', 'N', 'H', 'W', 'C', 'cstride', 'coff', 'lo')

    def __init__(self, buf, N, H, W, C, cstride=None, coff=0, lo=None):
        self.buf, self.N, self.H, self.W, self.C = buf, N, H, W, C
        self.cstride = C if cstride is None else cstride
        self.coff = coff
        self.lo = lo                         # 'fp32x3': buffer of the lo plane (same geometry), value = buf + lo

    def slice(self, off, c):
        return View(self.buf, self.N, self.H, self.W, c, self.cstride, self.coff + off, self.lo)

    def tensor(self):
        t = self.buf.view(self.N, self.H, self.W, self.cstride)[..., self.coff:self.coff + self.C]
        if self.lo is not None:
            t = t.float() + self.lo.view(self.N, self.H, self.W, self.cstride)[..., self.coff:self.coff + self.C].float()
        return t


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def _pool_out(h, k, s, pad, ceil_mode):
    if ceil_mode:
        o = int(math.ceil((h + 2 * pad - k) / float(s))) + 1
        if (o - 1) * s >= h + pad:
            o -= 1
    else:
        o = (h + 2 * pad - k) // s + 1
    return o


class Engine(object):
    def __init__(self, net, batch, precision, device, use_graph=True, trace=False):
        """``trace``: keep, per conv op, what a checker needs to recompute it from the op's own input buffer (folded fp32
        weights, geometry, the views it reads / writes) in ``self.trace`` — tests walk the compiled net layer by layer."""
        self.trace = [] if trace else None
        if precision not in _DT:
            raise ValueError("precision must be 'fp32', 'fp32x3', 'bf16' or 'fp16'")
        if not torch.cuda.is_available():
            raise _lib.CtxError('no CUDA device: the detection hot path has no CPU fallback')
        self.L = _lib.lib()
        self.dev = torch.device(device)
        if self.dev.type != 'cuda':
            raise _lib.CtxError('Engine needs a CUDA device, got %s' % device)
        if self.dev.index is None:
            self.dev = torch.device('cuda', torch.cuda.current_device())
        self.net_version = self._version_of(net)
        self.batch = batch
        self.precision = precision
        self.act_dtype = _DT[precision]
        self.act_code = _lib.dtype_code(self.act_dtype)
        self.keep = []                       # tensors whose raw pointers the program holds
        self.layers = []                     # (name, kind, flops) per op, for the benchmark / ncu tables
        self.prog = C.c_void_p()
        _lib.check(self.L.ctx_prog_create(C.byref(self.prog)), 'ctx_prog_create')
        self.graph_ready = False
        self.x_u8 = None
        self.rgb_means = tuple(float(m) for m in getattr(net, 'rgb_means', (104.0, 117.0, 123.0)))   # test.py:87
        self.use_graph = use_graph
        # independent chains (RFB branches, per-level heads) on their own graph lanes; CTX_LANES=0 keeps one chain
        self.use_lanes = os.environ.get('CTX_LANES', '1') != '0'
        self.autotune = os.environ.get('CTX_AUTOTUNE', '1') != '0' and precision in _TC16
        self.split = precision == 'fp32x3'
        self.stream = torch.cuda.Stream(device=self.dev)
        with torch.cuda.device(self.dev), torch.no_grad():
            self._compile(net)

    def __del__(self):
        try:
            if self.prog:
                self.L.ctx_prog_destroy(self.prog)
                self.prog = C.c_void_p()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _version_of(net):
        return tuple((t.data_ptr(), t._version) for t in list(net.parameters()) + list(net.buffers()))

    def stale(self, net):
        return self._version_of(net) != self.net_version

    def _alloc(self, *shape, dtype=None):
        t = torch.empty(*shape, dtype=dtype or self.act_dtype, device=self.dev)
        self.keep.append(t)
        return t

    def _lane(self, lane, wait=()):
        """Ops emitted from here on belong to ``lane``; the next one first waits for the lanes in ``wait``."""
        mask = 0
        for l in wait:
            mask |= 1 << l
        if not self.use_lanes:
            lane, mask = 0, 0
        _lib.check(self.L.ctx_prog_set_lane(self.prog, lane, mask), 'ctx_prog_set_lane')

    def _new_view(self, N, H, W, C):
        return View(self._alloc(N * H * W * C), N, H, W, C, lo=self._alloc(N * H * W * C) if self.split else None)

    # ------------------------------------------------------------------------------------------
    def _fold(self, conv, bn, scale=1.0):
        w = conv.weight.detach().to(self.dev, torch.float32)
        if bn is not None:
            s = bn.weight.detach().to(self.dev, torch.float32) / torch.sqrt(
                bn.running_var.detach().to(self.dev, torch.float32) + bn.eps)
            b = bn.bias.detach().to(self.dev, torch.float32) - bn.running_mean.detach().to(self.dev, torch.float32) * s
            w = w * s.view(-1, 1, 1, 1)
        elif conv.bias is not None:
            b = conv.bias.detach().to(self.dev, torch.float32)
        else:
            b = torch.zeros(w.size(0), device=self.dev)
        if scale != 1.0:
            w, b = w * scale, b * scale
        return w, b

    def _emit_conv(self, name, src, w, b, stride, pad, dil, relu, segs=None, out=None, residual=None, relu_channels=0,
                   algo_flops=None, in_nchw=False, pool2=False, stem=None):
        """w: folded fp32 [Cout,Cin,KH,KW]; b: fp32 [Cout].  segs: list of (tensor, c_begin, c_end, img_stride,
        pix_stride, ch_offset) for multi-destination epilogues; otherwise writes ``out`` (a View) or a new one.
        ``stem`` = (raw input View, w1, b1): this is conv1_2 and conv1_1 (3 -> 64, folded fp32 weights ``w1`` / bias ``b1``) is
        evaluated inside the same kernel from the raw fp32 NCHW input (csrc/conv_stem2.cu); ``src`` then only carries the
        geometry of the activation between the two convs, which never exists in HBM.  Returns None when the fused kernel does
        not take the geometry."""
        Cout, Cin, KH, KW = w.shape
        assert Cin == src.C, (name, Cin, src.C)
        ph, pw = pad
        if (src.H, src.W) == (1, 1) and stride == 1 and (KH, KW) == (3, 3) and (ph, pw) == (dil, dil) and not pool2 and stem is None:
            # a "same" 3x3 conv on a 1x1 map (the last head, RFB_Net_vgg.py:277-286 on the 1x1 source): eight of the nine taps only
            # ever see padding.  The centre tap as a 1x1 conv is the same sum bit for bit (the other products are exact zeros)
            # with a ninth of the weights to stream — this launch sits at the very end of the pyramid, in front of the attention.
            w = w[:, :, 1:2, 1:2].contiguous()
            KH = KW = 1
            ph = pw = 0
            dil = 1
        Ho = (src.H + 2 * ph - dil * (KH - 1) - 1) // stride + 1
        Wo = (src.W + 2 * pw - dil * (KW - 1) - 1) // stride + 1
        p = _lib.CtxConvParams()
        p.N, p.H, p.W, p.Cin = src.N, src.H, src.W, Cin
        p.in_cstride, p.in_coffset = src.cstride, src.coff
        p.Cout, p.KH, p.KW, p.stride, p.pad_h, p.pad_w, p.dil = Cout, KH, KW, stride, ph, pw, dil
        p.Ho, p.Wo, p.relu, p.relu_channels = Ho, Wo, int(relu), int(relu_channels)
        p.pool2 = int(pool2)
        p.in_dtype = _lib.dtype_code(src.buf.dtype)
        p.in_nchw = int(in_nchw)
        setattr(p, 'in', src.buf.data_ptr())
        bias = b.contiguous()
        self.keep.append(bias)
        p.bias = bias.data_ptr()
        if residual is not None:
            assert (residual.N, residual.H, residual.W, residual.C) == (src.N, Ho, Wo, Cout)
            p.residual = residual.buf.data_ptr()
            p.res_dtype = _lib.dtype_code(residual.buf.dtype)
            p.res_cstride, p.res_coffset = residual.cstride, residual.coff
        result = None
        if segs is None:
            Hq, Wq = (Ho // 2, Wo // 2) if pool2 else (Ho, Wo)
            if out is None:
                out = self._new_view(src.N, Hq, Wq, Cout)
            assert (out.N, out.H, out.W, out.C) == (src.N, Hq, Wq, Cout), name
            Ho_seg, Wo_seg = Hq, Wq
            segs = [(out.buf, 0, Cout, Ho_seg * Wo_seg * out.cstride, out.cstride, out.coff)]
            result = out
        p.nseg = len(segs)
        for i, (t, c0, c1, img_stride, pix_stride, ch_off) in enumerate(segs):
            p.seg[i].ptr = t.data_ptr()
            p.seg[i].c_begin, p.seg[i].c_end = c0, c1
            p.seg[i].img_stride, p.seg[i].pix_stride, p.seg[i].ch_offset = img_stride, pix_stride, ch_off
            p.seg[i].dtype = _lib.dtype_code(t.dtype)
        self.last_conv_params = p            # (tests re-plan the same conv under other tilings)
        if getattr(self, 'trace', None) is not None:
            self.trace.append(dict(name=name, src=src, w=w, b=b, stride=stride, pad=(ph, pw), dil=dil, relu=bool(relu),
                                   relu_channels=int(relu_channels), residual=residual, segs=segs, out=result, pool2=bool(pool2),
                                   in_nchw=bool(in_nchw), op=self.L.ctx_prog_num_ops(self.prog),
                                   stem=None if stem is None else dict(w=stem[1], b=stem[2])))
        flops = algo_flops if algo_flops is not None else 2.0 * src.N * Ho * Wo * Cout * Cin * KH * KW
        if self.split:
            return self._emit_conv_x3(name, p, src, w, residual, segs, result, flops, in_nchw, pool2)
        use_tc = self.precision != 'fp32' and self.L.ctx_conv2d_tc_supported(C.byref(p)) == 1
        if stem is not None:
            x_raw, w1, b1 = stem
            if not use_tc or self.L.ctx_conv2d_stem2_supported(C.byref(p)) != 1:
                if getattr(self, 'trace', None) is not None:
                    self.trace.pop()
                return None
            wt1 = torch.zeros(64, 64, dtype=self.act_dtype, device=self.dev)            # k = (ky*3 + kx)*3 + ci, as the STEM mode
            wt1[:, :27] = w1.permute(0, 2, 3, 1).reshape(64, 27).to(self.act_dtype)
            wt = torch.zeros((Cout + 15) // 16 * 16, KH * KW, 64, dtype=self.act_dtype, device=self.dev)
            wt[:Cout] = w.permute(0, 2, 3, 1).reshape(Cout, KH * KW, Cin).to(self.act_dtype)
            bias1 = b1.contiguous()
            self.keep += [wt1, wt, bias1]
            p.weight = wt.data_ptr()
            _lib.check(self.L.ctx_prog_add_conv_stem2(self.prog, C.byref(p), x_raw.buf.data_ptr(), wt1.data_ptr(), bias1.data_ptr()),
                       'ctx_prog_add_conv_stem2(%s)' % name)
            flops += 2.0 * src.N * src.H * src.W * 64 * 27
            self.layers.append((name, 'conv_tc', flops, (src.N, src.H, src.W, Cin, Cout, KH, KW, stride, dil)))
            return result
        if pool2 and not use_tc:
            return None                     # caller falls back to conv + separate pool
        if in_nchw and not use_tc:
            raise _lib.CtxError('stem conv: tensor-core STEM mode unavailable for this geometry')
        if use_tc:
            cin_p = (Cin + 63) // 64 * 64
            cout_p = (Cout + 15) // 16 * 16
            if in_nchw:                 # stem: one 64-wide K-step, k = (ky*3 + kx)*3 + ci
                wt = torch.zeros(cout_p, 64, dtype=self.act_dtype, device=self.dev)
                wt[:Cout, :27] = w.permute(0, 2, 3, 1).reshape(Cout, 27).to(self.act_dtype)
            else:
                wt = torch.zeros(cout_p, KH * KW, cin_p, dtype=self.act_dtype, device=self.dev)
                wt[:Cout, :, :Cin] = w.permute(0, 2, 3, 1).reshape(Cout, KH * KW, Cin).to(self.act_dtype)
            self.keep.append(wt)
            p.weight = wt.data_ptr()
            _lib.check(self.L.ctx_prog_add_conv_tc(self.prog, C.byref(p)), 'ctx_prog_add_conv_tc(%s)' % name)
            kind = 'conv_tc'
        else:
            cout_p = (Cout + 3) // 4 * 4
            wt = torch.zeros(KH * KW * Cin, cout_p, dtype=torch.float32, device=self.dev)
            wt[:, :Cout] = w.permute(2, 3, 1, 0).reshape(KH * KW * Cin, Cout)
            self.keep.append(wt)
            p.weight = wt.data_ptr()
            _lib.check(self.L.ctx_prog_add_conv_simt(self.prog, C.byref(p)), 'ctx_prog_add_conv_simt(%s)' % name)
            kind = 'conv_simt'
        self.layers.append((name, kind, flops, (src.N, src.H, src.W, Cin, Cout, KH, KW, stride, dil)))
        return result

    def _emit_conv_x3(self, name, p, src, w, residual, segs, result, flops, in_nchw, pool2):
        """'fp32x3' (csrc/conv_x3.cu): fp16 hi / lo planes of the activations and of the weights, the latter scaled per
        output channel by a power of two so that their lo plane stays in fp16's normal range (undone by ``out_scale``)."""
        if pool2:
            return None                     # the split-aware pool kernel follows as its own op
        Cout, Cin, KH, KW = w.shape
        amax = w.abs().amax(dim=(1, 2, 3)).clamp_min(1e-30)
        sc = torch.exp2(torch.floor(torch.log2(4096.0 / amax)))
        ws = w * sc.view(-1, 1, 1, 1)
        hi = ws.to(torch.float16)
        lo = (ws - hi.float()).to(torch.float16)
        cout_p = (Cout + 15) // 16 * 16
        if in_nchw:                         # stem: one 64-wide K-step per plane, k = (ky*3 + kx)*3 + ci
            wt = torch.zeros(cout_p, 2, 64, dtype=torch.float16, device=self.dev)
            wt[:Cout, 0, :27] = hi.permute(0, 2, 3, 1).reshape(Cout, 27)
            wt[:Cout, 1, :27] = lo.permute(0, 2, 3, 1).reshape(Cout, 27)
        else:
            cin_p = (Cin + 63) // 64 * 64
            wt = torch.zeros(cout_p, 2, KH * KW, cin_p, dtype=torch.float16, device=self.dev)
            wt[:Cout, 0, :, :Cin] = hi.permute(0, 2, 3, 1).reshape(Cout, KH * KW, Cin)
            wt[:Cout, 1, :, :Cin] = lo.permute(0, 2, 3, 1).reshape(Cout, KH * KW, Cin)
        inv = (1.0 / sc).contiguous()
        self.keep += [wt, inv]
        p.weight, p.out_scale, p.split = wt.data_ptr(), inv.data_ptr(), 1
        if not in_nchw:
            p.in_lo = src.lo.data_ptr()
        if residual is not None:
            p.residual_lo = residual.lo.data_ptr()
        if result is not None:
            p.out_lo = result.lo.data_ptr()
        _lib.check(self.L.ctx_prog_add_conv_x3(self.prog, C.byref(p)), 'ctx_prog_add_conv_x3(%s)' % name)
        self.layers.append((name, 'conv_x3', flops, (src.N, src.H, src.W, Cin, Cout, KH, KW, p.stride, p.dil)))
        return result

    def _emit_pool(self, name, src, k, s, pad, ceil_mode, out=None, in_img_stride=None, out_img_stride=None):
        Ho, Wo = _pool_out(src.H, k, s, pad, ceil_mode), _pool_out(src.W, k, s, pad, ceil_mode)
        if out is None:
            out = self._new_view(src.N, Ho, Wo, src.C)
        p = _lib.CtxPoolParams()
        p.N, p.H, p.W, p.C, p.Ho, p.Wo, p.k, p.stride, p.pad = src.N, src.H, src.W, src.C, Ho, Wo, k, s, pad
        p.dtype = _lib.dtype_code(src.buf.dtype)
        esz = src.buf.element_size()
        setattr(p, 'in', src.buf.data_ptr() + src.coff * esz)
        p.in_img_stride = in_img_stride if in_img_stride is not None else src.H * src.W * src.cstride
        p.in_pix_stride = src.cstride
        p.out = out.buf.data_ptr() + out.coff * out.buf.element_size()
        p.out_img_stride = out_img_stride if out_img_stride is not None else Ho * Wo * out.cstride
        p.out_pix_stride = out.cstride
        if src.lo is not None:
            p.in_lo = src.lo.data_ptr() + src.coff * esz
            p.out_lo = out.lo.data_ptr() + out.coff * esz
        _lib.check(self.L.ctx_prog_add_pool(self.prog, C.byref(p)), 'ctx_prog_add_pool(%s)' % name)
        self.layers.append((name, 'pool', 0.0, (src.N, src.H, src.W, src.C, k, s)))
        return out

    def _basic_conv(self, name, m, src, out=None, residual=None, relu=None, scale=1.0):
        w, b = self._fold(m.conv, m.bn, scale)
        c = m.conv
        return self._emit_conv(name, src, w, b, c.stride[0], _pair(c.padding), c.dilation[0],
                               (m.relu is not None) if relu is None else relu, out=out, residual=residual)

    def _rfb(self, name, m, src, home=0):
        """``home``: lane of the block's own chain (entry conv, first branch, ConvLinear).  The block input must already
        be complete on lane 0; with home != 0 the whole block runs beside the trunk (its output only feeds a head)."""
        branches = m.branches()
        if home:
            self._lane(home, wait=(0,))
        stride = m.shortcut.conv.stride[0]
        Ho = (src.H - 1) // stride + 1
        Wo = (src.W - 1) // stride + 1
        cat_c = sum(br[-1].out_channels for br in branches)
        cat = self._new_view(src.N, Ho, Wo, cat_c)

        # 1x1 convs that read the block input (branch entries + shortcut) are evaluated as ONE conv per
        # stride: output channels = [entries (ReLU) ..., shortcut (no ReLU)], later layers read channel slices.
        def is_pointwise(bc):
            c = bc.conv
            return c.kernel_size == (1, 1) and c.padding == (0, 0) and c.dilation == (1, 1)
        members = {}                                            # stride -> [(key, BasicConv)]
        for bi, br in enumerate(branches):
            if len(br) > 1 and is_pointwise(br[0]) and br[0].relu is not None:
                members.setdefault(br[0].conv.stride[0], []).append((bi, br[0]))
        if is_pointwise(m.shortcut) and m.shortcut.relu is None:
            members.setdefault(stride, []).append(('shortcut', m.shortcut))
        entry_out = {}
        for st, group in sorted(members.items()):
            if len(group) < 2 or any(bc.out_channels % 64 for _, bc in group):
                continue
            ws, bs = zip(*[self._fold(bc.conv, bc.bn) for _, bc in group])
            relu_c = sum(bc.out_channels for key, bc in group if key != 'shortcut')
            fused = self._emit_conv('%s.entry_s%d[%s]' % (name, st, '+'.join(str(k) for k, _ in group)), src,
                                    torch.cat(ws, 0), torch.cat(bs, 0), st, (0, 0), 1, relu_c > 0, relu_channels=relu_c)
            off = 0
            for key, bc in group:
                entry_out[key] = fused.slice(off, bc.out_channels)
                off += bc.out_channels

        # branch bi runs on lane bi (the block input / fused entry conv was produced on the home lane).  The side branches are
        # emitted FIRST: a lane's fork point is wherever the home lane stands when the lane's first op is recorded, so emitting
        # branch 0 (home lane) before them would make every side branch wait for branch 0's convs as well.
        offs, off = [], 0
        for br in branches:
            offs.append(off)
            off += br[-1].out_channels
        for bi in list(range(1, len(branches))) + [0]:
            br = branches[bi]
            self._lane(home if bi == 0 else min(bi, 3), wait=(home,))
            t = src
            for li, layer in enumerate(br):
                last = li == len(br) - 1
                if li == 0 and bi in entry_out:
                    t = entry_out[bi]
                    continue
                t = self._basic_conv('%s.branch%d.%d' % (name, bi, li), layer, t,
                                     out=cat.slice(offs[bi], layer.out_channels) if last else None)
        self._lane(home, wait=tuple(range(1, min(len(branches), 4))))
        short = entry_out['shortcut'] if 'shortcut' in entry_out else self._basic_conv(name + '.shortcut', m.shortcut, src)
        # relu(ConvLinear(cat) * scale + short): scale folds into the weights, the add + ReLU into the epilogue
        return self._basic_conv(name + '.ConvLinear', m.ConvLinear, cat, residual=short, relu=True, scale=float(m.scale))

    def _stem_pair(self, net, x_raw, w1, b1, relu1, hi):
        """conv1_1 -> ReLU -> conv1_2 -> ReLU [-> MaxPool2d(2,2)] (vgg() base.0 .. base.4, RFB_Net_vgg.py:323-343) as one kernel
        in the 16-bit modes: returns (output view, index of the next base module) or None.  CTX_STEM2=0 keeps the two convs apart."""
        if self.precision not in _TC16 or os.environ.get('CTX_STEM2', '1') == '0' or not relu1 or hi < 4:
            return None
        m = net.base[2]
        if not (isinstance(m, nn.Conv2d) and isinstance(net.base[3], nn.ReLU) and w1.size(0) == 64 and m.in_channels == 64 and m.out_channels <= 64
                and m.kernel_size == (3, 3) and m.stride == (1, 1) and m.padding == (1, 1) and m.dilation == (1, 1)):
            return None
        w2, b2 = self._fold(m, None)
        pool = net.base[4] if hi > 4 else None
        pool2 = (isinstance(pool, nn.MaxPool2d) and pool.kernel_size == 2 and pool.stride == 2 and pool.padding == 0
                 and x_raw.H % 2 == 0 and x_raw.W % 2 == 0)
        mid = View(torch.empty(0, dtype=self.act_dtype, device=self.dev), x_raw.N, x_raw.H, x_raw.W, 64)     # geometry only
        out = self._emit_conv('base.0+base.2' + ('+pool4' if pool2 else ''), mid, w2, b2, 1, (1, 1), 1, True, pool2=pool2,
                              stem=(x_raw, w1, b1))
        if out is None:
            return None
        return out, (5 if pool2 else 4)

    # ------------------------------------------------------------------------------------------
    def _compile(self, net):
        B, S = self.batch, net.size
        self.x_in = self._alloc(B, 3, S, S, dtype=torch.float32)
        stem = net.base[0]
        stem_as_gemm = (self.precision != 'fp32' and isinstance(stem, nn.Conv2d) and stem.in_channels == 3
                        and stem.kernel_size == (3, 3) and stem.stride == (1, 1) and stem.padding == (1, 1)
                        and stem.dilation == (1, 1))
        if stem_as_gemm:
            # Cin = 3 cannot feed the tensor cores channel-wise: the conv kernel's STEM mode reads the raw NCHW
            # fp32 input and builds each pixel's 3x3x3 patch (K = 27 -> one 64-wide K-step) on the fly
            x = View(self.x_in.view(-1), B, S, S, 3)
        else:
            x = self._new_view(B, S, S, 3)
            _lib.check(self.L.ctx_prog_add_nchw_to_nhwc(self.prog, self.x_in.data_ptr(), x.buf.data_ptr(), B, 3, S, S,
                                                        self.act_code), 'ctx_prog_add_nchw_to_nhwc')
            self.layers.append(('input.nhwc', 'layout', 0.0, (B, 3, S, S)))
        sources = []

        def run_base(lo, hi, x):
            k = lo
            while k < hi:
                m = net.base[k]
                if isinstance(m, nn.Conv2d):
                    relu = k + 1 < len(net.base) and isinstance(net.base[k + 1], nn.ReLU)
                    w, b = self._fold(m, None)
                    if k == 0 and stem_as_gemm:
                        fused = self._stem_pair(net, x, w, b, relu, hi)
                        if fused is not None:
                            x, k = fused
                            continue
                        x = self._emit_conv('base.0', x, w, b, 1, (1, 1), 1, relu, in_nchw=True)
                        k += 2 if relu else 1
                        continue
                    nxt = k + (2 if relu else 1)
                    pool = net.base[nxt] if nxt < hi else None
                    fused = None
                    if (self.precision in _TC16 and isinstance(pool, nn.MaxPool2d) and pool.kernel_size == 2 and pool.stride == 2
                            and pool.padding == 0 and x.H % 2 == 0 and x.W % 2 == 0 and m.stride[0] == 1):
                        # conv -> ReLU -> MaxPool2d(2,2): the pooled map is all that leaves the conv's epilogue
                        fused = self._emit_conv('base.%d+pool%d' % (k, nxt), x, w, b, 1, _pair(m.padding), m.dilation[0], relu, pool2=True)
                    if fused is not None:
                        x = fused
                        k = nxt + 1
                        continue
                    x = self._emit_conv('base.%d' % k, x, w, b, m.stride[0], _pair(m.padding), m.dilation[0], relu)
                    k += 2 if relu else 1
                elif isinstance(m, nn.MaxPool2d):
                    x = self._emit_pool('base.%d' % k, x, m.kernel_size, m.stride, m.padding, m.ceil_mode)
                    k += 1
                elif isinstance(m, nn.ReLU):
                    raise _lib.CtxError('base.%d: ReLU without a preceding conv' % k)
                else:
                    raise _lib.CtxError('base.%d: unsupported module %s' % (k, type(m).__name__))
            return x

        # ---- heads: one conv per level, three output segments, emitted as soon as the level's source exists and put
        # on lane 5 so that they overlap the rest of the trunk (they only meet again at the Context-Transformer) ------
        from .config import MBOX, VOC_300, VOC_512
        fmaps = (VOC_512 if S == 512 else VOC_300)['feature_maps']
        Csrc = net.num_classes
        anchors = [l.out_channels // 4 for l in net.loc]
        assert anchors == MBOX[S], (anchors, MBOX[S])
        level_p = [f * f * a for f, a in zip(fmaps, anchors)]
        P = sum(level_p)
        self.num_priors = P
        ours = net.ours
        if ours and len(fmaps) > len(CONF_POOL):
            raise IndexError('Context-Transformer pooling is defined for 6 pyramid levels only (size 300); size %d '
                             'has %d (undefined upstream as well, RFB_Net_vgg.py:235-243)' % (S, len(fmaps)))
        self.loc = self._alloc(B, P, 4, dtype=torch.float32)
        self.conf_raw = self._alloc(B, P, Csrc, dtype=torch.float32)
        self.obj_raw = self._alloc(B, P, 2, dtype=torch.float32)
        # conf max-pool (RFB_Net_vgg.py:242-244): level i's pooled map only needs head i, so it follows it on the heads' lane
        pooled_shapes = []
        if ours:
            for f, a, k in zip(fmaps, anchors, CONF_POOL):
                pooled_shapes.append((_pool_out(f, k, k, 0, True), _pool_out(f, k, k, 0, True), a))
            Pk = sum(h * w * a for h, w, a in pooled_shapes)
            self.num_pooled = Pk
            self.pooled = self._alloc(B, Pk, Csrc, dtype=torch.float32)

        def add_source(s, lane=0):
            i = len(sources)
            sources.append(s)
            assert (s.H, s.W) == (fmaps[i], fmaps[i]), (i, s.H, s.W, fmaps[i])
            a = anchors[i]
            poff = sum(level_p[:i])
            wl, bl = self._fold(net.loc[i], None)
            wc, bc = self._fold(net.conf[i], None)
            wo, bo = self._fold(net.obj[i], None)
            w = torch.cat([wl, wc, wo], 0)
            b = torch.cat([bl, bc, bo], 0)
            c1, c2, c3 = a * 4, a * 4 + a * Csrc, a * 4 + a * Csrc + a * 2
            segs = [(self.loc.view(-1)[poff * 4:], 0, c1, P * 4, a * 4, 0),
                    (self.conf_raw.view(-1)[poff * Csrc:], c1, c2, P * Csrc, a * Csrc, 0),
                    (self.obj_raw.view(-1)[poff * 2:], c2, c3, P * 2, a * 2, 0)]
            self._lane(5, wait=(lane,))
            self._emit_conv('head.%d' % i, s, w, b, 1, (1, 1), 1, False, segs=segs)
            if ours:
                hp, wp, _ = pooled_shapes[i]
                koff = sum(h * w * aa for h, w, aa in pooled_shapes[:i])
                src = View(self.conf_raw.view(-1), B, s.H, s.W, a * Csrc, a * Csrc, poff * Csrc)
                dst = View(self.pooled.view(-1), B, hp, wp, a * Csrc, a * Csrc, koff * Csrc)
                self._emit_pool('conf_pool.%d' % i, src, CONF_POOL[i], CONF_POOL[i], 0, True, out=dst,
                                in_img_stride=P * Csrc, out_img_stride=Pk * Csrc)
            self._lane(0)

        x = run_base(0, SOURCE_SPLIT, x)
        # RFB-a on conv4_3 only feeds the first head: it runs on lane 4 beside the rest of the trunk
        norm_lane = 4 if self.use_lanes else 0
        add_source(self._rfb('Norm', net.Norm, x, home=norm_lane), lane=norm_lane)
        x = run_base(SOURCE_SPLIT, len(net.base), x)
        for k, m in enumerate(net.extras):
            if isinstance(m, _RFBBlock):
                x = self._rfb('extras.%d' % k, m, x)
            elif isinstance(m, BasicConv):
                x = self._basic_conv('extras.%d' % k, m, x)
            else:
                raise _lib.CtxError('extras.%d: unsupported module %s' % (k, type(m).__name__))
            if k < net.indicator or k % 2 == 0:
                add_source(x)
        assert len(sources) == len(fmaps)
        self._lane(0, wait=(5,))
        self.head_end_op = self.L.ctx_prog_num_ops(self.prog)

        # ---- Context-Transformer (phase 2, method 'ours') ----------------------------------------
        if ours:
            incre = net.setting == 'incre'
            n_novel = net.OBJ_Target.out_features
            n_out = n_novel + (Csrc if incre else 0)
            self.conf = self._alloc(B, P, n_out, dtype=torch.float32)
            f32 = lambda t: self._hold(t.detach().to(self.dev, torch.float32).contiguous())
            ap = _lib.CtxAttnParams()
            ap.batch, ap.num_priors, ap.num_pooled, ap.dim = B, P, Pk, Csrc
            ap.num_novel, ap.incre, ap.apply_softmax = n_novel, int(incre), 1
            ap.conf, ap.pooled = self.conf_raw.data_ptr(), self.pooled.data_ptr()
            ap.theta_w, ap.theta_b = f32(net.theta.weight), f32(net.theta.bias)
            ap.phi_w, ap.phi_b = f32(net.phi.weight), f32(net.phi.bias)
            ap.g_w, ap.g_b = f32(net.g.weight), f32(net.g.bias)
            if incre:
                ap.fc_base_w, ap.fc_base_b = f32(net.fc_base.weight), f32(net.fc_base.bias)
            ap.Wz = f32(net.Wz)
            ap.obj_target_w = f32(net.OBJ_Target.weight)
            ap.scale = float(net.scale.detach().float().cpu().item())
            # 0: fp32 CUDA cores; 1: tcgen05, fp16 logits; 2: tcgen05, fp16 hi/lo split logits; 3: + hi/lo split P and V
            ap.use_tensor_cores = {'fp32': 0, 'fp32x3': 3}.get(self.precision, 1)
            if os.environ.get('CTX_ATTN_MODE'):                  # development aid: force a Context-Transformer kernel variant
                ap.use_tensor_cores = int(os.environ['CTX_ATTN_MODE'])
            ws_bytes = self.L.ctx_attention_workspace_bytes(C.byref(ap))
            self.attn_ws = self._alloc(ws_bytes + 1024, dtype=torch.uint8)
            ws_ptr = (self.attn_ws.data_ptr() + 1023) // 1024 * 1024
            ap.workspace, ap.workspace_bytes, ap.out = ws_ptr, ws_bytes, self.conf.data_ptr()
            _lib.check(self.L.ctx_prog_add_attention(self.prog, C.byref(ap)), 'ctx_prog_add_attention')
            self.layers.append(('context_transformer', 'attention', 4.0 * B * P * Pk * Csrc, (B, P, Pk, Csrc)))
        else:
            self.conf = self._alloc(B, P, Csrc, dtype=torch.float32)
            _lib.check(self.L.ctx_prog_add_softmax(self.prog, self.conf_raw.data_ptr(), self.conf.data_ptr(), B * P, Csrc),
                       'ctx_prog_add_softmax')
            self.layers.append(('conf.softmax', 'softmax', 0.0, (B * P, Csrc)))
        self.obj = self._alloc(B, P, 2, dtype=torch.float32)
        self._lane(5)                        # independent of the conf path: runs beside the Context-Transformer kernel
        _lib.check(self.L.ctx_prog_add_softmax(self.prog, self.obj_raw.data_ptr(), self.obj.data_ptr(), B * P, 2),
                   'ctx_prog_add_softmax')
        self.layers.append(('obj.softmax', 'softmax', 0.0, (B * P, 2)))
        self._lane(0, wait=(5,))
        self.num_ops = self.L.ctx_prog_num_ops(self.prog)
        if self.autotune:
            # per-layer tiling picked by timing each candidate on this GPU (bit-identical outputs)
            torch.cuda.synchronize(self.dev)
            _lib.check(self.L.ctx_prog_autotune(self.prog, C.c_void_p(self.stream.cuda_stream), 4), 'ctx_prog_autotune')
        self.conv_flops = sum(l[2] for l in self.layers if l[1].startswith('conv'))

    def _hold(self, t):
        self.keep.append(t)
        return t.data_ptr()

    # ------------------------------------------------------------------------------------------
    def load_input(self, x):
        """Stage ``x`` ([B,3,S,S], any device, fp32) into the program's input buffer (async on the
        current stream; a pinned host tensor becomes one H2D copy).  A uint8 ``[B,S,S,3]`` batch (cv2 images that already
        have the network's size) takes the on-device ``BaseTransform`` path instead: a quarter of the H2D bytes, the
        mean subtraction and HWC -> CHW change done by ``ctx_base_transform``."""
        if x.dtype == torch.uint8:
            B, _, S, _ = self.x_in.shape
            if tuple(x.shape) != (B, S, S, 3):
                raise ValueError('uint8 input must be [B,S,S,3] = %s (resize on the host first), got %s' % ((B, S, S, 3), tuple(x.shape)))
            if self.x_u8 is None:
                self.x_u8 = torch.empty(B, S, S, 3, dtype=torch.uint8, device=self.dev)
            self.x_u8.copy_(x, non_blocking=True)
            means = (C.c_float * 3)(*self.rgb_means)
            _lib.check(self.L.ctx_base_transform(self.x_u8.data_ptr(), self.x_in.data_ptr(), B, S, S, means,
                                                 _lib.current_stream_ptr(self.dev)), 'ctx_base_transform')
            return
        if tuple(x.shape) != tuple(self.x_in.shape):
            raise ValueError('engine compiled for input %s, got %s' % (tuple(self.x_in.shape), tuple(x.shape)))
        self.x_in.copy_(x, non_blocking=True)

    def launch(self):
        """Replay the program on the current stream (input already staged)."""
        st = torch.cuda.current_stream(self.dev)
        if self.use_graph and not self.graph_ready:
            # capture on the engine's private stream, then replay on the caller's stream
            self.stream.wait_stream(st)
            with torch.cuda.stream(self.stream):
                _lib.check(self.L.ctx_prog_run_range(self.prog, 0, self.num_ops, C.c_void_p(self.stream.cuda_stream)),
                           'ctx_prog_run (warm-up)')
                self.stream.synchronize()
                _lib.check(self.L.ctx_prog_instantiate_graph(self.prog, C.c_void_p(self.stream.cuda_stream)),
                           'ctx_prog_instantiate_graph')
            st.wait_stream(self.stream)
            self.graph_ready = True
        _lib.check(self.L.ctx_prog_run(self.prog, C.c_void_p(st.cuda_stream)), 'ctx_prog_run')

    def run(self, x):
        with torch.cuda.device(self.dev):
            self.load_input(x)
            self.launch()
        return self.loc, self.conf, self.obj

    def conv_config(self, op_index):
        """(tile width, N tiles, cluster, A mode, stages, grid, commit group, patch) of a tensor-core conv op, zeros otherwise."""
        info = (C.c_int * 8)()
        _lib.check(self.L.ctx_prog_conv_config(self.prog, op_index, info), 'ctx_prog_conv_config')
        return list(info)

    def run_range(self, first, last, reps=1):
        with torch.cuda.device(self.dev):
            st = torch.cuda.current_stream(self.dev)
            _lib.check(self.L.ctx_prog_run_range_repeat(self.prog, first, last, reps, C.c_void_p(st.cuda_stream)), 'ctx_prog_run_range')

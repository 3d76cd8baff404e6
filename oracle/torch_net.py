"""TEST INFRASTRUCTURE ONLY — plain torch fp32 (CPU) functional restatement of the reference
detector forward, driven directly by a ``state_dict`` (no nn.Module, no product code).

Follows reference models/RFB_Net_vgg.py: BasicConv :7-22, BasicRFB :26-64, BasicRFB_a :68-112,
RFBNet.forward :190-286 (Context-Transformer :253-271), vgg() :323-343, add_extras :354-378,
multibox :387-416.  Used as the floating-point reference for the CUDA conv / pool / attention
kernels; pinned against the real reference by tests/golden/net_*.npz (oracle/gen_golden.py).
"""
import torch
import torch.nn.functional as F

VGG_CFG = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'C', 512, 512, 512, 'M', 512, 512, 512]
POOL_K = [3, 2, 2, 2, 1, 1]          # RFB_Net_vgg.py:235-236 (6 entries: 512 + 'ours' is undefined upstream)


class Quant(object):
    """Quantisation-faithful variant of the forward (``forward(..., quant=Quant(torch.bfloat16))``): the SAME graph with the
    roundings the 16-bit tensor-core engine performs made explicit, so that its kernels can be held to a tight tolerance
    against an oracle that shares their quantisation instead of a loose one against fp32:
      * BatchNorm folded into the conv weights (w * gamma / sqrt(var + eps)) BEFORE they are rounded to 16 bits; the
        folded bias stays fp32; accumulation in fp32;
      * every activation rounded to 16 bits once, where a kernel's epilogue stores it (after bias [+ shortcut] [ReLU]);
        the network input is rounded by the stem; head outputs (loc / conf / obj) stay fp32;
      * Context-Transformer (csrc/attention_tc.cu): K, V and the probabilities P rounded to fp16, Q too when
        ``split_logits`` is False (use_tensor_cores = 1; = 2 keeps hi + lo parts: fp32-grade logits)."""

    def __init__(self, act_dtype, split_logits=False):
        self.dt = act_dtype
        self.split_logits = split_logits

    def a(self, x):
        return x.to(self.dt).float()

    @staticmethod
    def h(x):
        return x.half().float()


def _fold(sd, name):
    s = sd[name + '.bn.weight'] / torch.sqrt(sd[name + '.bn.running_var'] + 1e-5)
    return sd[name + '.conv.weight'] * s.view(-1, 1, 1, 1), sd[name + '.bn.bias'] - sd[name + '.bn.running_mean'] * s


def _basic_conv(sd, name, x, stride=1, padding=0, dilation=1, relu=True, q=None, residual=None):
    """conv(bias=False) -> BN(eval, eps 1e-5) -> optional ReLU   (BasicConv :7-22).  ``residual`` (quantised path only):
    added before the ReLU, the way the engine's ConvLinear epilogue applies the RFB shortcut (:59-61)."""
    if q is not None:
        w, b = _fold(sd, name)
        y = F.conv2d(x, q.a(w), None, stride, padding, dilation) + b.view(1, -1, 1, 1)
        if residual is not None:
            y = y + residual
        return q.a(F.relu(y) if relu else y)
    x = F.conv2d(x, sd[name + '.conv.weight'], None, stride, padding, dilation)
    x = F.batch_norm(x, sd[name + '.bn.running_mean'], sd[name + '.bn.running_var'],
                     sd[name + '.bn.weight'], sd[name + '.bn.bias'], False, 0.0, 1e-5)
    return F.relu(x) if relu else x


def _vgg_conv(sd, k, x, padding=1, dilation=1, q=None):
    """conv with bias + ReLU of the VGG trunk (vgg() :323-343)"""
    w, b = sd['base.%d.weight' % k], sd['base.%d.bias' % k]
    if q is not None:
        return q.a(F.relu(F.conv2d(x, q.a(w), None, 1, padding, dilation) + b.view(1, -1, 1, 1)))
    return F.relu(F.conv2d(x, w, b, 1, padding, dilation))


def _rfb_a(sd, p, x, q=None):
    """BasicRFB_a(512,512,stride=1,scale=1.0)  (:68-112)"""
    b0 = _basic_conv(sd, p + '.branch0.0', x, q=q)
    b0 = _basic_conv(sd, p + '.branch0.1', b0, padding=1, relu=False, q=q)
    b1 = _basic_conv(sd, p + '.branch1.0', x, q=q)
    b1 = _basic_conv(sd, p + '.branch1.1', b1, padding=(1, 0), q=q)
    b1 = _basic_conv(sd, p + '.branch1.2', b1, padding=3, dilation=3, relu=False, q=q)
    b2 = _basic_conv(sd, p + '.branch2.0', x, q=q)
    b2 = _basic_conv(sd, p + '.branch2.1', b2, padding=(0, 1), q=q)
    b2 = _basic_conv(sd, p + '.branch2.2', b2, padding=3, dilation=3, relu=False, q=q)
    b3 = _basic_conv(sd, p + '.branch3.0', x, q=q)
    b3 = _basic_conv(sd, p + '.branch3.1', b3, padding=(0, 1), q=q)
    b3 = _basic_conv(sd, p + '.branch3.2', b3, padding=(1, 0), q=q)
    b3 = _basic_conv(sd, p + '.branch3.3', b3, padding=5, dilation=5, relu=False, q=q)
    short = _basic_conv(sd, p + '.shortcut', x, relu=False, q=q)
    if q is not None:
        return _basic_conv(sd, p + '.ConvLinear', torch.cat((b0, b1, b2, b3), 1), relu=True, q=q, residual=short)
    out = _basic_conv(sd, p + '.ConvLinear', torch.cat((b0, b1, b2, b3), 1), relu=False)
    return F.relu(out * 1.0 + short)


def _rfb(sd, p, x, stride, visual, q=None):
    """BasicRFB(in,out,stride,scale=1.0,visual)  (:26-64)"""
    b0 = _basic_conv(sd, p + '.branch0.0', x, stride=stride, q=q)
    b0 = _basic_conv(sd, p + '.branch0.1', b0, padding=visual, dilation=visual, relu=False, q=q)
    b1 = _basic_conv(sd, p + '.branch1.0', x, q=q)
    b1 = _basic_conv(sd, p + '.branch1.1', b1, stride=stride, padding=1, q=q)
    b1 = _basic_conv(sd, p + '.branch1.2', b1, padding=visual + 1, dilation=visual + 1, relu=False, q=q)
    b2 = _basic_conv(sd, p + '.branch2.0', x, q=q)
    b2 = _basic_conv(sd, p + '.branch2.1', b2, padding=1, q=q)
    b2 = _basic_conv(sd, p + '.branch2.2', b2, stride=stride, padding=1, q=q)
    b2 = _basic_conv(sd, p + '.branch2.3', b2, padding=2 * visual + 1, dilation=2 * visual + 1, relu=False, q=q)
    short = _basic_conv(sd, p + '.shortcut', x, stride=stride, relu=False, q=q)
    if q is not None:
        return _basic_conv(sd, p + '.ConvLinear', torch.cat((b0, b1, b2), 1), relu=True, q=q, residual=short)
    out = _basic_conv(sd, p + '.ConvLinear', torch.cat((b0, b1, b2), 1), relu=False)
    return F.relu(out * 1.0 + short)


def extras_spec(size):
    """[(kind, stride, visual | (k, pad))] per extras index (add_extras :354-378)."""
    if size == 300:
        return [('rfb', 1, 2), ('rfb', 2, 2), ('rfb', 2, 2),
                ('conv', 1, 0), ('conv', 3, 0), ('conv', 1, 0), ('conv', 3, 0)]
    return [('rfb', 1, 2), ('rfb', 2, 2), ('rfb', 2, 2), ('rfb', 2, 1), ('rfb', 2, 1),
            ('conv', 1, 0), ('conv', 4, 1)]


def backbone_sources(sd, x, size, q=None):
    """sources list (conv4_3->Norm, then the extras taps)  (forward :218-233)"""
    k = 0
    if q is not None:
        x = q.a(x)                               # the stem's producer warps round the fp32 image to 16 bits
    for v in VGG_CFG[:13]:                       # up to conv4_3 + ReLU == base[0..22]
        if v == 'M':
            x = F.max_pool2d(x, 2, 2)
        elif v == 'C':
            x = F.max_pool2d(x, 2, 2, ceil_mode=True)
        else:
            x = _vgg_conv(sd, k, x, q=q)
            k += 1
        k += 1
    sources = [_rfb_a(sd, 'Norm', x, q)]
    for v in VGG_CFG[13:]:
        if v == 'M':
            x = F.max_pool2d(x, 2, 2)
        else:
            x = _vgg_conv(sd, k, x, q=q)
            k += 1
        k += 1
    x = F.max_pool2d(x, 3, 1, 1)                                                   # pool5 == base[30]
    x = _vgg_conv(sd, 31, x, 6, 6, q)                                              # conv6
    x = _vgg_conv(sd, 33, x, 0, 1, q)                                              # conv7
    indicator = 3 if size == 300 else 5
    for i, (kind, a, b) in enumerate(extras_spec(size)):
        if kind == 'rfb':
            x = _rfb(sd, 'extras.%d' % i, x, a, b, q)
        else:
            x = _basic_conv(sd, 'extras.%d' % i, x, padding=b, q=q)
        if i < indicator or i % 2 == 0:
            sources.append(x)
    return sources


def forward(sd, x, size, num_classes, method='ours', phase=2, setting='transfer', training=False,
            init=False, return_parts=False, quant=None):
    """Returns (loc[B,P,4], conf[B,P,C'], obj[B,P,2]) exactly as RFBNet.forward (:190-286).  ``quant``: see ``Quant``."""
    q = quant
    sd = {k: v.detach().float() for k, v in sd.items()}
    x = x.float()
    num = x.size(0)
    ours = (method == 'ours' and phase == 2)
    sources = backbone_sources(sd, x, size, q)
    loc, conf, obj, conf_pool = [], [], [], []
    hw = (lambda w: q.a(w)) if q is not None else (lambda w: w)            # head weights are 16-bit operands too; outputs stay fp32
    for i, s in enumerate(sources):
        loc.append(F.conv2d(s, hw(sd['loc.%d.weight' % i]), sd['loc.%d.bias' % i], 1, 1).permute(0, 2, 3, 1).reshape(num, -1))
        c = F.conv2d(s, hw(sd['conf.%d.weight' % i]), sd['conf.%d.bias' % i], 1, 1)
        conf.append(c.permute(0, 2, 3, 1).reshape(num, -1))
        obj.append(F.conv2d(s, hw(sd['obj.%d.weight' % i]), sd['obj.%d.bias' % i], 1, 1).permute(0, 2, 3, 1).reshape(num, -1))
        if ours:
            conf_pool.append(F.max_pool2d(c, POOL_K[i], POOL_K[i], ceil_mode=True).permute(0, 2, 3, 1).reshape(num, -1))
    loc = torch.cat(loc, 1)
    conf = torch.cat(conf, 1)
    obj = torch.cat(obj, 1)
    if init:
        return conf.view(num, -1, num_classes)
    parts = {}
    if ours:
        conf_pool = torch.cat(conf_pool, 1).view(num, -1, num_classes)
        conf = conf.view(num, -1, num_classes)
        if setting == 'incre':
            conf_base = F.linear(conf, sd['fc_base.weight'], sd['fc_base.bias']) + conf
        q = F.linear(conf, sd['theta.weight'], sd['theta.bias']) + conf
        kk = F.linear(conf_pool, sd['phi.weight'], sd['phi.bias']) + conf_pool
        v = F.linear(conf_pool, sd['g.weight'], sd['g.bias']) + conf_pool
        if quant is not None:
            # csrc/attention_tc.cu: fp16 K / V / P operands (Q too unless the logits use the hi/lo split), fp32 accumulate;
            # p = exp(s - rowmax) un-normalised in fp16, row sum accumulated from the same rounded values
            ql = q if quant.split_logits else Quant.h(q)
            kl = kk if quant.split_logits else Quant.h(kk)
            s_ = torch.matmul(ql, kl.transpose(1, 2))
            pr = Quant.h(torch.exp(s_ - s_.max(dim=2, keepdim=True).values))
            delta = torch.matmul(pr, Quant.h(v)) / pr.sum(dim=2, keepdim=True) * sd['Wz']
        else:
            w = torch.softmax(torch.matmul(q, kk.transpose(1, 2)), dim=2)
            delta = torch.matmul(w, v) * sd['Wz']
        z = conf + delta
        z = z / z.norm(dim=2, keepdim=True)
        novel = F.linear(z, sd['OBJ_Target.weight']) * sd['scale']
        parts = dict(conf_raw=conf, conf_pool=conf_pool, q=q, k=kk, v=v, z=z)
        conf = novel if setting == 'transfer' else torch.cat((conf_base, novel), dim=2)
    else:
        conf = conf.view(num, -1, num_classes)
    loc = loc.view(num, -1, 4)
    obj = obj.view(num, -1, 2)
    if not training:
        conf = torch.softmax(conf, dim=-1)
        obj = torch.softmax(obj, dim=-1)
    if return_parts:
        return (loc, conf, obj), parts
    return loc, conf, obj

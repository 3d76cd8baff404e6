"""``RFBNet`` / ``build_net`` — the detector module, drop-in for reference ``models/RFB_Net_vgg.py``.

Same call surface (``build_net(args, size, num_classes)`` -> ``nn.Module`` with ``forward(x,
init=False)``, ``normalize()``, ``load_weights()``, caller-assigned ``device``, attribute ``size``)
and the same ``state_dict()`` keys and shapes (``base.N.weight``, ``Norm.branchK.J.conv.weight``,
``extras.N...``, ``loc/conf/obj.N``, ``theta/phi/g``, ``Wz``, ``scale``, ``OBJ_Target``[,
``fc_base``]) so reference checkpoints, ``utils/solver.py`` (LR by parameter name) and
``train.py:284-286`` keep working.

What differs is how it runs.  The layers are described once by a small table (``_vgg_plan``,
``_rfb_plan`` ...); the ``nn.Module`` tree only owns the parameters.  In ``eval()`` mode the whole
forward — layout change, every conv with BatchNorm / bias / ReLU / residual folded into its
epilogue, pools, the three heads of a level as ONE conv writing straight into the concatenated
``[B,P,*]`` buffers, conf max-pool, the Context-Transformer block and the output softmaxes — is
compiled by ``engine.Engine`` into a list of hand-written sm_100a kernels and replayed with one C
call.  There is no PyTorch fallback for inference: without a GPU or without ``libctx_b200.so`` the
forward raises.  In ``train()`` mode (and for ``init=True`` prototype extraction, train.py:252-286)
the forward is expressed with autograd tensor ops on the same parameters, because the fine-tune
loop needs gradients through training-mode BatchNorm (SURVEY.md §7 "hard parts"); only its loss
kernels (``match`` / mining) are native.
"""
import os
from collections import namedtuple

import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.nn.init as init

ConvPlan = namedtuple('ConvPlan', 'cin cout k stride pad dil relu')

VGG_PLAN = (64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'C', 512, 512, 512, 'M', 512, 512, 512)
MBOX = {300: (6, 6, 6, 6, 4, 4), 512: (6, 6, 6, 6, 6, 4, 4)}
SOURCE_SPLIT = 23                      # base[0:23] ends at conv4_3 + ReLU (RFB_Net_vgg.py:219-220)
CONF_POOL = (3, 2, 2, 2, 1, 1)         # kernel == stride of the conf max-pool per level (:235-236)


def _cp(cin, cout, k, stride=1, pad=0, dil=1, relu=True):
    return ConvPlan(cin, cout, k, stride, pad, dil, relu)


class BasicConv(nn.Module):
    """conv(bias=False) -> BatchNorm(eps 1e-5, momentum 0.01) -> ReLU; bn / relu optional."""

    def __init__(self, in_planes, out_planes, kernel_size, stride=1, padding=0, dilation=1, groups=1, relu=True,
                 bn=True, bias=False):
        super(BasicConv, self).__init__()
        if groups != 1:
            raise ValueError('BasicConv: grouped convolutions are not part of this network')
        self.out_channels = out_planes
        self.conv = nn.Conv2d(in_planes, out_planes, kernel_size=kernel_size, stride=stride, padding=padding,
                              dilation=dilation, bias=bias)
        self.bn = nn.BatchNorm2d(out_planes, eps=1e-5, momentum=0.01, affine=True) if bn else None
        self.relu = nn.ReLU(inplace=True) if relu else None

    def forward(self, x):
        x = self.conv(x)
        if self.bn is not None:
            x = self.bn(x)
        return x if self.relu is None else self.relu(x)


def _branch(plans):
    return nn.Sequential(*[BasicConv(p.cin, p.cout, p.k, stride=p.stride, padding=p.pad, dilation=p.dil, relu=p.relu)
                           for p in plans])


class _RFBBlock(nn.Module):
    """Multi-branch dilated block: cat(branches) -> 1x1 ConvLinear, *scale + 1x1 shortcut, ReLU."""

    def __init__(self, branch_plans, in_planes, out_planes, stride, scale):
        super(_RFBBlock, self).__init__()
        self.scale = scale
        self.out_channels = out_planes
        self.num_branches = len(branch_plans)
        for i, plans in enumerate(branch_plans):
            setattr(self, 'branch%d' % i, _branch(plans))
        cat_planes = sum(plans[-1].cout for plans in branch_plans)
        self.ConvLinear = BasicConv(cat_planes, out_planes, kernel_size=1, stride=1, relu=False)
        self.shortcut = BasicConv(in_planes, out_planes, kernel_size=1, stride=stride, relu=False)
        self.relu = nn.ReLU(inplace=True)

    def branches(self):
        return [getattr(self, 'branch%d' % i) for i in range(self.num_branches)]

    def forward(self, x):
        out = self.ConvLinear(torch.cat([b(x) for b in self.branches()], 1))
        return self.relu(out * self.scale + self.shortcut(x))


class BasicRFB(_RFBBlock):
    """RFB (reference RFB_Net_vgg.py:26-64): three branches with dilations visual, visual+1, 2*visual+1."""

    def __init__(self, in_planes, out_planes, stride=1, scale=0.1, visual=1):
        i = in_planes // 8
        plans = [
            [_cp(in_planes, 2 * i, 1, stride), _cp(2 * i, 2 * i, 3, 1, visual, visual, relu=False)],
            [_cp(in_planes, i, 1), _cp(i, 2 * i, 3, stride, 1), _cp(2 * i, 2 * i, 3, 1, visual + 1, visual + 1, relu=False)],
            [_cp(in_planes, i, 1), _cp(i, (i // 2) * 3, 3, 1, 1), _cp((i // 2) * 3, 2 * i, 3, stride, 1),
             _cp(2 * i, 2 * i, 3, 1, 2 * visual + 1, 2 * visual + 1, relu=False)],
        ]
        super(BasicRFB, self).__init__(plans, in_planes, out_planes, stride, scale)


class BasicRFB_a(_RFBBlock):
    """RFB-s (reference RFB_Net_vgg.py:68-112): four branches with (3,1)/(1,3) factorised kernels."""

    def __init__(self, in_planes, out_planes, stride=1, scale=0.1):
        i = in_planes // 4
        plans = [
            [_cp(in_planes, i, 1), _cp(i, i, 3, 1, 1, relu=False)],
            [_cp(in_planes, i, 1), _cp(i, i, (3, 1), 1, (1, 0)), _cp(i, i, 3, 1, 3, 3, relu=False)],
            [_cp(in_planes, i, 1), _cp(i, i, (1, 3), stride, (0, 1)), _cp(i, i, 3, 1, 3, 3, relu=False)],
            [_cp(in_planes, i // 2, 1), _cp(i // 2, (i // 4) * 3, (1, 3), 1, (0, 1)),
             _cp((i // 4) * 3, i, (3, 1), stride, (1, 0)), _cp(i, i, 3, 1, 5, 5, relu=False)],
        ]
        super(BasicRFB_a, self).__init__(plans, in_planes, out_planes, stride, scale)


def vgg(cfg, i, batch_norm=False):
    """VGG16 trunk with fc6/fc7 as dilated conv6 / 1x1 conv7 (reference :323-343)."""
    layers, cin = [], i
    for v in cfg:
        if v == 'M' or v == 'C':
            layers.append(nn.MaxPool2d(kernel_size=2, stride=2, ceil_mode=(v == 'C')))
            continue
        layers.append(nn.Conv2d(cin, v, kernel_size=3, padding=1))
        if batch_norm:
            layers.append(nn.BatchNorm2d(v))
        layers.append(nn.ReLU(inplace=True))
        cin = v
    layers += [nn.MaxPool2d(kernel_size=3, stride=1, padding=1),
               nn.Conv2d(512, 1024, kernel_size=3, padding=6, dilation=6), nn.ReLU(inplace=True),
               nn.Conv2d(1024, 1024, kernel_size=1), nn.ReLU(inplace=True)]
    return layers


base = {'300': list(VGG_PLAN), '512': list(VGG_PLAN)}
extras = {'300': [1024, 'S', 512, 'S', 256], '512': [1024, 'S', 512, 'S', 256, 'S', 256, 'S', 256]}
mbox = {'300': list(MBOX[300]), '512': list(MBOX[512])}


def add_extras(size, cfg, in_channels):
    """RFB pyramid blocks + the plain tail convs (reference :354-378)."""
    if size not in (300, 512):
        print("Error: Sorry only RFBNet300 and RFBNet512 are supported!")
        return None
    layers, cin = [], in_channels
    for k, v in enumerate(cfg):
        if cin != 'S':
            if v == 'S':
                layers.append(BasicRFB(cin, cfg[k + 1], stride=2, scale=1.0, visual=1 if (cin == 256 and size == 512) else 2))
            else:
                layers.append(BasicRFB(cin, v, scale=1.0, visual=2))
        cin = v
    layers.append(BasicConv(256, 128, kernel_size=1, stride=1))
    if size == 512:
        layers.append(BasicConv(128, 256, kernel_size=4, stride=1, padding=1))
    else:
        layers += [BasicConv(128, 256, kernel_size=3, stride=1), BasicConv(256, 128, kernel_size=1, stride=1),
                   BasicConv(128, 256, kernel_size=3, stride=1)]
    return layers


def source_indices(size, n_extras):
    indicator = 3 if size == 300 else 5
    return [k for k in range(n_extras) if k < indicator or k % 2 == 0]


def multibox(size, vgg_layers, extra_layers, cfg, num_classes):
    """Per pyramid level: 3x3 loc (A*4), conf (A*num_classes), obj (A*2) heads (reference :387-416)."""
    if size not in (300, 512):
        print("Error: Sorry only RFBNet300 and RFBNet512 are supported!")
        return None
    chans = [512] + [extra_layers[k].out_channels for k in source_indices(size, len(extra_layers))]
    heads = ([], [], [])
    for c, a in zip(chans, cfg):
        for lst, per in zip(heads, (4, num_classes, 2)):
            lst.append(nn.Conv2d(c, a * per, kernel_size=3, padding=1))
    return vgg_layers, extra_layers, heads


class RFBNet(nn.Module):
    """precision: arithmetic of the compiled inference path — 'fp32' (CUDA cores, 1e-4 parity
    mode) or 'bf16' / 'fp16' (tcgen05 tensor cores, fp32 accumulate).  Set ``net.precision``
    before the first eval forward or call ``net.invalidate_engine()`` after changing it."""

    def __init__(self, args, size, base, extras, head, num_classes):
        super(RFBNet, self).__init__()
        self.method = args.method
        self.phase = args.phase
        self.setting = args.setting
        self.num_classes = num_classes
        self.size = size
        self.precision = getattr(args, 'precision', 'fp32')
        self.use_cuda_graph = True
        self.static_outputs = False        # True: forward returns the engine's own output buffers (zero-copy; overwritten by the next call)
        self._engines = {}
        if size == 300:
            self.indicator = 3
        elif size == 512:
            self.indicator = 5
        else:
            print("Error: Sorry only SSD300 and SSD512 are supported!")
            return
        self.base = nn.ModuleList(base)
        self.Norm = BasicRFB_a(512, 512, stride=1, scale=1.0)
        self.extras = nn.ModuleList(extras)
        self.loc = nn.ModuleList(head[0])
        self.conf = nn.ModuleList(head[1])
        self.obj = nn.ModuleList(head[2])
        self.init_weight()
        if self.ours:
            d, n_novel = (60, 20) if args.setting == 'transfer' else (15, 5)
            if args.setting == 'incre':
                self.fc_base = nn.Linear(d, d)
                self.fc_base.weight.data.fill_(0)
                self.fc_base.bias.data.fill_(0)
            elif args.setting != 'transfer':
                return
            self.theta = nn.Linear(d, d)
            self.phi = nn.Linear(d, d)
            self.g = nn.Linear(d, d)
            self.Wz = nn.Parameter(torch.zeros(d))
            self.OBJ_Target = nn.Linear(d, n_novel, bias=False)
            self.scale = nn.Parameter(torch.FloatTensor([5]), requires_grad=False)
            for lin in (self.theta, self.phi, self.g):
                init.kaiming_normal_(lin.weight, mode='fan_out')
                lin.bias.data.fill_(0)

    @property
    def ours(self):
        return self.method == 'ours' and self.phase == 2

    # ---- inference: compiled sm_100a program --------------------------------------------------
    MAX_ENGINES = 3                        # live compiled programs (one per batch size / precision), least recently used goes first

    def invalidate_engine(self):
        self._engines = {}

    def engine(self, batch):
        """The compiled program for this batch size (activations are sized per batch).  A few are kept (the last, smaller
        batch of a dataset must not evict the main one); a parameter change (``load_state_dict``, ``normalize()``, an
        optimizer step) drops all of them.  Re-tuning after such a rebuild is free: the per-layer tiling found by
        measurement is cached per conv geometry inside the library."""
        from .engine import Engine
        dev = torch.device(self.device)
        key = (batch, self.precision, dev.index if dev.index is not None else torch.cuda.current_device(), self.use_cuda_graph)
        eng = self._engines.pop(key, None)
        if eng is not None and eng.stale(self):
            self._engines, eng = {}, None
        if eng is None:
            eng = Engine(self, batch, self.precision, dev, use_graph=self.use_cuda_graph)
        self._engines[key] = eng                   # most recently used last
        while len(self._engines) > self.MAX_ENGINES:
            self._engines.pop(next(iter(self._engines)))
        return eng

    def forward(self, x, init=False):
        if self.training or init:
            return self._forward_autograd(x, init)
        if not hasattr(self, 'device'):
            raise AttributeError("RFBNet.device is not set (the caller assigns model.device = 'cuda', test.py:191-196)")
        if torch.device(self.device).type != 'cuda':
            raise RuntimeError("RFBNet inference runs only on a CUDA device (sm_100a kernels, no CPU fallback); "
                               "got model.device = %r" % (self.device,))
        out = self.engine(x.size(0)).run(x)            # fp32 [B,3,S,S], or uint8 [B,S,S,3] (on-device BaseTransform)
        # like the reference (RFB_Net_vgg.py:246-286) every call returns tensors of its own: the caller may keep them while
        # the next batch runs.  ``static_outputs = True`` hands out the engine's buffers instead (saves three D2D copies).
        return out if self.static_outputs else tuple(t.clone() for t in out)

    # ---- training / prototype init: autograd expression of the same graph ---------------------
    def _forward_autograd(self, x, init=False):
        x = x.to(self.device)
        num = x.size(0)
        sources = []
        for k in range(SOURCE_SPLIT):
            x = self.base[k](x)
        sources.append(self.Norm(x))
        for k in range(SOURCE_SPLIT, len(self.base)):
            x = self.base[k](x)
        for k, v in enumerate(self.extras):
            x = v(x)
            if k < self.indicator or k % 2 == 0:
                sources.append(x)
        if self.ours and len(sources) > len(CONF_POOL):
            raise IndexError('Context-Transformer pooling is defined for 6 pyramid levels only (size 300); '
                             'size %d has %d (undefined upstream as well, RFB_Net_vgg.py:235-243)' % (self.size, len(sources)))
        loc, conf, obj, conf_pool = [], [], [], []
        for i, (s, l, c, o) in enumerate(zip(sources, self.loc, self.conf, self.obj)):
            loc.append(l(s).permute(0, 2, 3, 1).reshape(num, -1))
            cmap = c(s)                                   # computed once (the reference evaluates it twice)
            conf.append(cmap.permute(0, 2, 3, 1).reshape(num, -1))
            obj.append(o(s).permute(0, 2, 3, 1).reshape(num, -1))
            if self.ours:
                conf_pool.append(F.max_pool2d(cmap, CONF_POOL[i], CONF_POOL[i], ceil_mode=True)
                                 .permute(0, 2, 3, 1).reshape(num, -1))
        loc, conf, obj = torch.cat(loc, 1), torch.cat(conf, 1), torch.cat(obj, 1)
        if init:
            return conf.view(num, -1, self.num_classes)
        if self.ours:
            conf = conf.view(num, -1, self.num_classes)
            pool = torch.cat(conf_pool, 1).view(num, -1, self.num_classes)
            q = self.theta(conf) + conf
            k_ = self.phi(pool) + pool
            v_ = self.g(pool) + pool
            attn = F.softmax(torch.matmul(q, k_.transpose(1, 2)), dim=2)
            z = conf + torch.matmul(attn, v_) * self.Wz
            z = z / z.norm(dim=2, keepdim=True)
            novel = self.OBJ_Target(z) * self.scale
            conf = novel if self.setting == 'transfer' else torch.cat((self.fc_base(conf) + conf, novel), dim=2)
        else:
            conf = conf.view(num, -1, self.num_classes)
        loc, obj = loc.view(num, -1, 4), obj.view(num, -1, 2)
        if self.training:
            return loc, conf, obj
        return loc, F.softmax(conf, dim=-1), F.softmax(obj, dim=-1)

    def load_weights(self, base_file):
        _, ext = os.path.splitext(base_file)
        if ext in ('.pkl', '.pth'):
            print('Loading weights into state dict...')
            self.load_state_dict(torch.load(base_file, map_location='cpu'))
            print('Finished!')
        else:
            print('Sorry only .pth and .pkl files supported.')

    def init_weight(self):
        """kaiming-normal (fan_out) on BasicConv convs, BN weight 1, every bias 0 (reference :297-314)."""
        for group in (self.base, self.Norm, self.extras, self.loc, self.conf, self.obj):
            for name, t in group.state_dict().items():
                leaf = name.split('.')[-1]
                if leaf == 'weight':
                    if 'conv' in name:
                        init.kaiming_normal_(t, mode='fan_out')
                    if 'bn' in name:
                        t.fill_(1)
                elif leaf == 'bias':
                    t.fill_(0)

    def normalize(self):
        self.OBJ_Target.weight.data = self.OBJ_Target.weight / self.OBJ_Target.weight.norm(dim=1, keepdim=True)


def build_net(args, size, num_classes):
    if size != 300 and size != 512:
        print("Error: Sorry only RFBNet300 and RFBNet512 are supported!")
        return
    return RFBNet(args, size, *multibox(size, vgg(base[str(size)], 3), add_extras(size, extras[str(size)], 1024),
                                        mbox[str(size)], num_classes), num_classes)

"""TEST INFRASTRUCTURE ONLY — seeded synthetic weights / inputs / head outputs / targets.

The reference ships no weights, data or golden vectors that are reachable offline (SURVEY.md §4,
§8c), so every parity test runs on these deterministic generators.  They use only torch's CPU
generator (bit-reproducible for a given torch build, same image here and on the GPU box) and are
keyed by parameter NAME, so the same values are produced whether the template state_dict comes
from the real reference module (gen_golden.py) or from the drop-in module (tests).
"""
import zlib

import numpy as np
import torch


def _gen(seed, key):
    g = torch.Generator()
    g.manual_seed((int(seed) * 1000003 + zlib.crc32(key.encode())) % (2 ** 63 - 1))
    return g


def seeded_state(template, seed=0, input_scale=40.0):
    """template: mapping name -> tensor (shapes/dtypes are used, values ignored).

    * conv weights  ~ N(0, 2/fan_in) (ReLU-preserving); ``base.0`` additionally /input_scale so the
      activations are O(1) for inputs of mean-subtracted-pixel magnitude;
    * head convs (loc/conf/obj) ~ N(0, 0.04/fan_in) -> loc / logits O(1) on the O(6) feature maps;
    * conv biases ~ N(0, 0.05²); BN affine/stats randomised away from identity;
    * theta/phi/g ~ N(0, 2/60), biases N(0, .05²); Wz ~ N(0, 0.5²) (zero-init upstream would make
      the attention branch a no-op); OBJ_Target rows L2-normalised (== normalize()); scale = 5.
    """
    out = {}
    for key in sorted(template.keys()):
        t = template[key]
        shape = tuple(t.shape)
        g = _gen(seed, key)
        leaf = key.split('.')[-1]
        if leaf == 'num_batches_tracked':
            out[key] = torch.zeros(shape, dtype=torch.long)
            continue
        if key == 'scale':
            out[key] = torch.full(shape, 5.0)
            continue
        if key == 'Wz':
            out[key] = 0.5 * torch.randn(shape, generator=g)
            continue
        if '.bn.' in key:
            if leaf == 'weight':
                v = 0.5 + torch.rand(shape, generator=g)
            elif leaf == 'running_var':
                v = 0.5 + torch.rand(shape, generator=g)
            else:                                   # bias, running_mean
                v = 0.1 * torch.randn(shape, generator=g)
            out[key] = v
            continue
        if leaf == 'bias':
            out[key] = 0.05 * torch.randn(shape, generator=g)
            continue
        # weights
        if len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            head = key.split('.')[0] in ('loc', 'conf', 'obj')
            std = 0.2 * (1.0 / fan_in) ** 0.5 if head else (2.0 / fan_in) ** 0.5
            if key == 'base.0.weight':
                std /= input_scale
            out[key] = std * torch.randn(shape, generator=g)
        elif key.startswith('OBJ_Target'):
            w = torch.randn(shape, generator=g)
            out[key] = w / w.norm(dim=1, keepdim=True)
        elif key.startswith('fc_base'):
            out[key] = (0.5 / shape[1]) ** 0.5 * torch.randn(shape, generator=g)
        else:                                       # theta / phi / g
            out[key] = (2.0 / shape[1]) ** 0.5 * torch.randn(shape, generator=g)
    return out


def seeded_input(batch, size, seed=0, scale=40.0):
    g = _gen(seed, 'input%dx%d' % (batch, size))
    return scale * torch.randn(batch, 3, size, size, generator=g)


def calibrated_heads(batch, num_priors, num_fg_classes=20, seed=0, pos_frac=0.02):
    """Head outputs that resemble a trained detector (SURVEY §8d "calibrated case"):
    ~pos_frac of priors are objects; their class logits get +6 on one random class.
    Returns loc[B,P,4] ~N(0,1), conf[B,P,C] softmax probs, obj[B,P,2] softmax probs (eval-mode
    outputs of the net, i.e. what Detect consumes)."""
    g = _gen(seed, 'heads%d_%d_%d' % (batch, num_priors, num_fg_classes))
    loc = torch.randn(batch, num_priors, 4, generator=g)
    conf_logit = torch.randn(batch, num_priors, num_fg_classes, generator=g)
    is_pos = torch.rand(batch, num_priors, generator=g) < pos_frac
    cls = torch.randint(0, num_fg_classes, (batch, num_priors), generator=g)
    boost = torch.zeros_like(conf_logit)
    boost.scatter_(2, cls.unsqueeze(-1), 6.0)
    conf_logit = conf_logit + boost * is_pos.unsqueeze(-1)
    obj_logit = torch.randn(batch, num_priors, 2, generator=g)
    obj_logit[..., 1] += torch.where(is_pos, torch.tensor(4.0), torch.tensor(-4.0))
    return loc, torch.softmax(conf_logit, -1), torch.softmax(obj_logit, -1)


def synthetic_targets(batch, seed=0, num_classes=20, max_obj=4):
    """Config-5 targets: list of [n,6] = x1,y1,x2,y2,label,weight (voc0712.py:246-248 format);
    n~U{1..max_obj}, xy~U(0,.5), wh~U(.1,.5), labels U{1..num_classes}, weight 1."""
    g = _gen(seed, 'targets%d' % batch)
    out = []
    for _ in range(batch):
        n = int(torch.randint(1, max_obj + 1, (1,), generator=g))
        xy = 0.5 * torch.rand(n, 2, generator=g)
        wh = 0.1 + 0.4 * torch.rand(n, 2, generator=g)
        lab = torch.randint(1, num_classes + 1, (n, 1), generator=g).float()
        out.append(torch.cat([xy, xy + wh, lab, torch.ones(n, 1)], 1))
    return out


def random_dets(n, seed=0, extent=(500.0, 375.0), tie_free=True):
    """n random pixel-space boxes with scores for NMS tests (float32 [n,5])."""
    rng = np.random.RandomState(seed)
    cx = rng.uniform(0, extent[0], n)
    cy = rng.uniform(0, extent[1], n)
    w = rng.uniform(10, 120, n)
    h = rng.uniform(10, 120, n)
    s = rng.uniform(0.01, 1.0, n)
    d = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2, s], 1).astype(np.float32)
    if tie_free:
        # make float32 scores distinct
        u, idx = np.unique(d[:, 4], return_index=True)
        if len(u) != n:
            d[:, 4] = (np.argsort(np.argsort(d[:, 4], kind='stable'), kind='stable').astype(np.float32) + 1) / (n + 1)
    return d

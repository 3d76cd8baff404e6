"""``PriorBox`` — SSD default boxes in centre form (host, one-time).

Mirror of reference ``layers/functions/prior_box.py:12-56``: same constructor contract (cfg keys,
``ValueError`` on a non-positive variance) and the same emission order (level → row i(y) →
column j(x) → [min, sqrt(min*max), then per aspect ratio (w*sqrt(ar), h/sqrt(ar)) and its
transpose]).  All arithmetic is Python float64, converted to float32 once at the end, exactly as
the reference does via ``torch.Tensor(list)``, so the result is bit-identical.
"""
from math import sqrt

import numpy as np
import torch


class PriorBox(object):
    def __init__(self, cfg):
        super(PriorBox, self).__init__()
        self.image_size = cfg['min_dim']
        self.num_priors = len(cfg['aspect_ratios'])
        self.variance = cfg['variance'] or [0.1]
        self.feature_maps = cfg['feature_maps']
        self.min_sizes = cfg['min_sizes']
        self.max_sizes = cfg['max_sizes']
        self.steps = cfg['steps']
        self.aspect_ratios = cfg['aspect_ratios']
        self.clip = cfg['clip']
        for v in self.variance:
            if v <= 0:
                raise ValueError('Variances must be greater than 0')

    def _level(self, k, f):
        f_k = self.image_size / self.steps[k]
        s_k = self.min_sizes[k] / self.image_size
        s_k_prime = sqrt(s_k * (self.max_sizes[k] / self.image_size))
        shapes = [(s_k, s_k), (s_k_prime, s_k_prime)]
        for ar in self.aspect_ratios[k]:
            r = sqrt(ar)
            shapes.append((s_k * r, s_k / r))
            shapes.append((s_k / r, s_k * r))
        rows = []
        for i in range(f):
            cy = (i + 0.5) / f_k
            for j in range(f):
                cx = (j + 0.5) / f_k
                for (w, h) in shapes:
                    rows.append((cx, cy, w, h))
        return rows

    def forward(self):
        rows = []
        for k, f in enumerate(self.feature_maps):
            rows.extend(self._level(k, f))
        output = torch.from_numpy(np.asarray(rows, dtype=np.float64).astype(np.float32)).view(-1, 4)
        if self.clip:
            output.clamp_(max=1, min=0)
        return output

"""Prior-box configuration constants of the detection path.

Restates the numeric cfg dicts the reference keeps in ``data/config.py`` (VOC_300 :10-26,
VOC_512 :46-62, COCO_300 :65-81, COCO_512 :101-117).  Only the constants the hot path reads
(``PriorBox`` / ``Detect`` / loss) are kept; dataset roots and unused SSD/mobile variants are not.
"""


def _rfb_cfg(min_dim, feature_maps, steps, min_sizes, max_sizes, aspect_ratios):
    return {
        'feature_maps': list(feature_maps),
        'min_dim': min_dim,
        'steps': list(steps),
        'min_sizes': list(min_sizes),
        'max_sizes': list(max_sizes),
        'aspect_ratios': [list(a) for a in aspect_ratios],
        'variance': [0.1, 0.2],
        'clip': True,
    }


_AR300 = ([2, 3], [2, 3], [2, 3], [2, 3], [2], [2])
_AR512 = ([2, 3], [2, 3], [2, 3], [2, 3], [2, 3], [2], [2])

VOC_300 = _rfb_cfg(300, (38, 19, 10, 5, 3, 1), (8, 16, 32, 64, 100, 300),
                   (30, 60, 111, 162, 213, 264), (60, 111, 162, 213, 264, 315), _AR300)
VOC_512 = _rfb_cfg(512, (64, 32, 16, 8, 4, 2, 1), (8, 16, 32, 64, 128, 256, 512),
                   (35.84, 76.8, 153.6, 230.4, 307.2, 384.0, 460.8),
                   (76.8, 153.6, 230.4, 307.2, 384.0, 460.8, 537.6), _AR512)
COCO_300 = _rfb_cfg(300, (38, 19, 10, 5, 3, 1), (8, 16, 32, 64, 100, 300),
                    (21, 45, 99, 153, 207, 261), (45, 99, 153, 207, 261, 315), _AR300)
COCO_512 = _rfb_cfg(512, (64, 32, 16, 8, 4, 2, 1), (8, 16, 32, 64, 128, 256, 512),
                    (20.48, 51.2, 133.12, 215.04, 296.96, 378.88, 460.8),
                    (51.2, 133.12, 215.04, 296.96, 378.88, 460.8, 542.72), _AR512)

# anchors per feature-map cell (reference models/RFB_Net_vgg.py:419-422)
MBOX = {300: [6, 6, 6, 6, 4, 4], 512: [6, 6, 6, 6, 6, 4, 4]}


def num_priors(cfg):
    """P for a cfg: sum over levels of f*f*(2 + 2*len(aspect_ratios))."""
    return sum(f * f * (2 + 2 * len(a)) for f, a in zip(cfg['feature_maps'], cfg['aspect_ratios']))

"""TEST INFRASTRUCTURE ONLY.  Import the real reference from /root/reference.

Exists only in the build container (the GPU box has no /root/reference); used by gen_golden.py
and by CPU tests that are skipped when the reference is absent.  Follows SURVEY.md Appendix B.
"""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get('CTX_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REF_ROOT, 'models', 'RFB_Net_vgg.py'))


_cache = {}


def load():
    """Return a namespace with the reference's hot-path symbols."""
    if 'ns' in _cache:
        return _cache['ns']
    if not available():
        raise RuntimeError('reference not present at %s' % REF_ROOT)
    sys.dont_write_bytecode = True          # /root/reference is read-only
    saved = list(sys.path)
    saved_mods = {k: sys.modules.get(k) for k in ('models', 'layers', 'utils', 'data')}
    for k in saved_mods:
        sys.modules.pop(k, None)
    sys.path.insert(0, REF_ROOT)
    try:
        from models.RFB_Net_vgg import build_net, RFBNet
        from layers.functions import Detect, PriorBox
        from layers.modules.multibox_loss_combined import MultiBoxLoss_combined
        from utils import box_utils
        spec = importlib.util.spec_from_file_location('refcfg', os.path.join(REF_ROOT, 'data', 'config.py'))
        cfg = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(cfg)
        spec = importlib.util.spec_from_file_location(
            'ref_py_cpu_nms', os.path.join(REF_ROOT, 'utils', 'nms', 'py_cpu_nms.py'))
        pynms = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(pynms)
    finally:
        sys.path[:] = saved
    ns = types.SimpleNamespace(build_net=build_net, RFBNet=RFBNet, Detect=Detect, PriorBox=PriorBox,
                               MultiBoxLoss_combined=MultiBoxLoss_combined, box_utils=box_utils,
                               cfg=cfg, py_cpu_nms=pynms.py_cpu_nms, cpu_nms=None, cpu_soft_nms=None)
    # optional: the reference's own Cython NMS, built by oracle/build.py into oracle/_ref/
    refdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')
    if os.path.isdir(refdir):
        sys.path.insert(0, refdir)
        try:
            import cpu_nms as _c
            ns.cpu_nms, ns.cpu_soft_nms = _c.cpu_nms, _c.cpu_soft_nms
        except Exception:
            pass
        finally:
            sys.path.remove(refdir)
    _cache['ns'] = ns
    return ns

"""``nms`` / ``soft_nms`` — numpy-in / list-out NMS with the reference's wrapper semantics.

Mirror of reference ``utils/nms_wrapper.py:23-31``: ``nms(dets[n,5] float32, thresh, force_cpu)``
returns the kept row indices, best score first, ``[]`` for empty input.  ``force_cpu`` does not
move the work to the CPU (there is no CPU path in this package): it selects the arithmetic
convention of the reference's CPU routine — suppress when ``ovr >= thresh`` (cpu_nms.pyx:65) —
while the default is the GPU routine's ``ovr > thresh`` (nms_kernel.cu:71, py_cpu_nms.py:36).
Rows with equal scores are ordered by row index (the reference's argsort is unstable there).

``gpu_nms`` / ``cpu_nms`` / ``cpu_soft_nms`` are provided under the reference's names
(gpu_nms.pyx:16-31, cpu_nms.pyx:17-68, :70-163).
"""
import ctypes as C

import numpy as np

from . import _lib


def _device_id(device_id):
    import torch
    if not torch.cuda.is_available():
        raise _lib.CtxError('nms: no CUDA device (the NMS of this package runs only on the GPU)')
    return torch.cuda.current_device() if device_id is None else int(device_id)


def _nms(dets, thresh, on_equal, device_id=None):
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    if dets.ndim != 2 or dets.shape[1] != 5:
        raise ValueError('nms: dets must be [n,5] (x1,y1,x2,y2,score)')
    n = dets.shape[0]
    if n == 0:
        return []
    keep = np.empty(n, dtype=np.int32)
    num = C.c_int(0)
    _lib.check(_lib.lib().ctx_nms_host(dets.ctypes.data, n, float(thresh), int(on_equal), keep.ctypes.data,
                                       C.addressof(num), _device_id(device_id)), 'ctx_nms_host')
    return keep[:num.value].tolist()


def gpu_nms(dets, thresh, device_id=None):
    return _nms(dets, thresh, False, device_id)


def cpu_nms(dets, thresh):
    return _nms(dets, thresh, True)


def nms(dets, thresh, force_cpu=False):
    """Dispatch exactly like the reference wrapper."""
    if dets.shape[0] == 0:
        return []
    if force_cpu:
        return cpu_nms(dets, thresh)
    return gpu_nms(dets, thresh)


def cpu_soft_nms(boxes, sigma=0.5, Nt=0.3, threshold=0.001, method=0, device_id=None):
    """In place on ``boxes`` [n,5] float32 (rows are reordered, scores decayed), returns
    ``list(range(N_final))`` — cpu_nms.pyx:70-163.  method: 0 hard, 1 linear, 2 gaussian."""
    if not (isinstance(boxes, np.ndarray) and boxes.dtype == np.float32 and boxes.flags['C_CONTIGUOUS']
            and boxes.ndim == 2 and boxes.shape[1] == 5):
        raise ValueError('cpu_soft_nms: boxes must be a C-contiguous float32 ndarray [n,5]')
    n = boxes.shape[0]
    if n == 0:
        return []
    n_out = C.c_int(0)
    _lib.check(_lib.lib().ctx_soft_nms_host(boxes.ctypes.data, n, float(sigma), float(Nt), float(threshold),
                                            int(method), C.addressof(n_out), _device_id(device_id)),
               'ctx_soft_nms_host')
    return list(range(n_out.value))


soft_nms = cpu_soft_nms


def nms_device(dets, thresh, suppress_on_equal=False):
    """Device-resident variant: ``dets`` CUDA float32 [n,5] -> CUDA int32 keep[num] (no host copies
    except the kept count)."""
    import torch
    _lib.require_cuda(dets, 'dets')
    dets = dets.float().contiguous()
    n = dets.size(0)
    if n == 0:
        return torch.empty(0, dtype=torch.int32, device=dets.device)
    L = _lib.lib()
    ws = torch.empty(L.ctx_nms_workspace_bytes(n), dtype=torch.uint8, device=dets.device)
    keep = torch.empty(n, dtype=torch.int32, device=dets.device)
    num = torch.zeros(1, dtype=torch.int32, device=dets.device)
    with torch.cuda.device(dets.device):
        _lib.check(L.ctx_nms_device(dets.data_ptr(), n, float(thresh), int(suppress_on_equal), keep.data_ptr(),
                                    num.data_ptr(), ws.data_ptr(), ws.numel(), _lib.current_stream_ptr()),
                   'ctx_nms_device')
    return keep[:int(num.item())]

"""ctypes binding of ``libctx_b200.so`` (declared in ``include/ctx_b200.h``).

The library is the product: there is no Python / CPU fallback.  ``lib()`` raises if the shared
object is missing (build it with ``python -m context_transformer_b200.build`` or
``__graft_entry__.build()``), and every wrapper raises ``CtxError`` on a non-zero status with the
library's thread-local message.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libctx_b200.so')

CTX_F32, CTX_BF16, CTX_F16 = 0, 1, 2
STATUS = {1: 'CTX_ERR_INVALID', 2: 'CTX_ERR_CUDA', 3: 'CTX_ERR_WORKSPACE', 4: 'CTX_ERR_UNSUPPORTED'}


class CtxError(RuntimeError):
    pass


class CtxPostParams(C.Structure):
    _fields_ = [('batch', C.c_int), ('num_priors', C.c_int), ('num_fg_classes', C.c_int),
                ('var0', C.c_float), ('var1', C.c_float), ('scale_per_image', C.c_int),
                ('score_thresh', C.c_float), ('nms_thresh', C.c_float), ('suppress_on_equal', C.c_int),
                ('nms_method', C.c_int), ('soft_sigma', C.c_float), ('soft_threshold', C.c_float),
                ('max_per_image', C.c_int), ('max_out', C.c_int)]


class CtxOutSeg(C.Structure):
    _fields_ = [('ptr', C.c_void_p), ('c_begin', C.c_int), ('c_end', C.c_int), ('img_stride', C.c_longlong),
                ('pix_stride', C.c_int), ('ch_offset', C.c_int), ('dtype', C.c_int)]


class CtxConvParams(C.Structure):
    _fields_ = [('N', C.c_int), ('H', C.c_int), ('W', C.c_int), ('Cin', C.c_int),
                ('in_cstride', C.c_int), ('in_coffset', C.c_int),
                ('Cout', C.c_int), ('KH', C.c_int), ('KW', C.c_int), ('stride', C.c_int),
                ('pad_h', C.c_int), ('pad_w', C.c_int), ('dil', C.c_int),
                ('Ho', C.c_int), ('Wo', C.c_int), ('relu', C.c_int), ('pool2', C.c_int), ('relu_channels', C.c_int), ('in_dtype', C.c_int), ('in_nchw', C.c_int),
                ('in', C.c_void_p), ('weight', C.c_void_p), ('bias', C.c_void_p), ('residual', C.c_void_p),
                ('res_dtype', C.c_int), ('res_cstride', C.c_int), ('res_coffset', C.c_int),
                ('nseg', C.c_int), ('seg', CtxOutSeg * 3),
                ('split', C.c_int), ('in_lo', C.c_void_p), ('residual_lo', C.c_void_p), ('out_lo', C.c_void_p),
                ('out_scale', C.c_void_p)]


class CtxPoolParams(C.Structure):
    _fields_ = [('N', C.c_int), ('H', C.c_int), ('W', C.c_int), ('C', C.c_int), ('Ho', C.c_int), ('Wo', C.c_int),
                ('k', C.c_int), ('stride', C.c_int), ('pad', C.c_int), ('dtype', C.c_int),
                ('in', C.c_void_p), ('in_img_stride', C.c_longlong), ('in_pix_stride', C.c_int),
                ('out', C.c_void_p), ('out_img_stride', C.c_longlong), ('out_pix_stride', C.c_int),
                ('in_lo', C.c_void_p), ('out_lo', C.c_void_p)]


class CtxAttnParams(C.Structure):
    _fields_ = [('batch', C.c_int), ('num_priors', C.c_int), ('num_pooled', C.c_int), ('dim', C.c_int),
                ('num_novel', C.c_int), ('incre', C.c_int), ('apply_softmax', C.c_int),
                ('conf', C.c_void_p), ('pooled', C.c_void_p),
                ('theta_w', C.c_void_p), ('theta_b', C.c_void_p), ('phi_w', C.c_void_p), ('phi_b', C.c_void_p),
                ('g_w', C.c_void_p), ('g_b', C.c_void_p), ('fc_base_w', C.c_void_p), ('fc_base_b', C.c_void_p),
                ('Wz', C.c_void_p), ('obj_target_w', C.c_void_p), ('scale', C.c_float),
                ('use_tensor_cores', C.c_int), ('workspace', C.c_void_p), ('workspace_bytes', C.c_size_t),
                ('out', C.c_void_p)]


_P, _I, _F, _SZ, _LL = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_longlong
# name -> (restype, argtypes); every symbol include/ctx_b200.h declares
SIGNATURES = {
    'ctx_version': (_I, []),
    'ctx_last_error': (C.c_char_p, []),
    'ctx_launch_count': (C.c_ulonglong, []),
    'ctx_detect_forward': (_I, [_P, _P, _P, _P, _I, _I, _I, _F, _F, _P, _P, _P]),
    'ctx_postprocess_workspace_bytes': (_SZ, [_I, _I, _I]),
    'ctx_detect_postprocess': (_I, [_P, _P, _P, _P, _P, C.POINTER(CtxPostParams), _P, _P, _P, _P, _SZ, _P]),
    'ctx_nms_workspace_bytes': (_SZ, [_I]),
    'ctx_nms_device': (_I, [_P, _I, _F, _I, _P, _P, _P, _SZ, _P]),
    'ctx_nms_host': (_I, [_P, _I, _F, _I, _P, _P, _I]),
    '_nms': (None, [_P, _P, _P, _I, _I, _F, _I]),
    'ctx_soft_nms_host': (_I, [_P, _I, _F, _F, _F, C.c_uint, _P, _I]),
    'ctx_conv2d_simt': (_I, [C.POINTER(CtxConvParams), _P]),
    'ctx_conv2d_tc_supported': (_I, [C.POINTER(CtxConvParams)]),
    'ctx_conv2d_tc_plan_create': (_I, [C.POINTER(CtxConvParams), C.POINTER(_P)]),
    'ctx_conv2d_tc_plan_create_tuned': (_I, [C.POINTER(CtxConvParams), _I, _I, _I, _I, C.POINTER(_P)]),
    'ctx_conv2d_tc_plan_info': (_I, [_P, C.POINTER(_I)]),
    'ctx_conv2d_tc_plan_run': (_I, [_P, _P]),
    'ctx_conv2d_tc_plan_destroy': (None, [_P]),
    'ctx_conv2d_x3_supported': (_I, [C.POINTER(CtxConvParams)]),
    'ctx_conv2d_x3_plan_create': (_I, [C.POINTER(CtxConvParams), _I, C.POINTER(_P)]),
    'ctx_conv2d_x3_plan_run': (_I, [_P, _P]),
    'ctx_conv2d_x3_plan_info': (_I, [_P, C.POINTER(_I)]),
    'ctx_conv2d_x3_plan_destroy': (None, [_P]),
    'ctx_prog_add_conv_x3': (_I, [_P, C.POINTER(CtxConvParams)]),
    'ctx_debug_set_conv_timeline': (None, [_P]),
    'ctx_conv2d_stem2_supported': (_I, [C.POINTER(CtxConvParams)]),
    'ctx_conv2d_stem2_plan_create': (_I, [C.POINTER(CtxConvParams), _P, _P, _P, C.POINTER(_P)]),
    'ctx_prog_add_conv_stem2': (_I, [_P, C.POINTER(CtxConvParams), _P, _P, _P]),
    'ctx_maxpool2d_nhwc': (_I, [C.POINTER(CtxPoolParams), _P]),
    'ctx_nchw_to_nhwc': (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    'ctx_base_transform': (_I, [_P, _P, _I, _I, _I, C.POINTER(C.c_float), _P]),
    'ctx_base_transform_resize': (_I, [_P, _I, _I, _P, _I, C.POINTER(C.c_float), _P]),
    'ctx_nchw_to_patch27': (_I, [_P, _P, _I, _I, _I, _I, _P]),
    'ctx_prog_add_nchw_to_patch27': (_I, [_P, _P, _P, _I, _I, _I, _I]),
    'ctx_attention_workspace_bytes': (_SZ, [C.POINTER(CtxAttnParams)]),
    'ctx_attention_forward': (_I, [C.POINTER(CtxAttnParams), _P]),
    'ctx_debug_set_attention_timeline': (None, [_P]),
    'ctx_softmax_lastdim': (_I, [_P, _P, _LL, _I, _P]),
    'ctx_prog_create': (_I, [C.POINTER(_P)]),
    'ctx_prog_add_conv_simt': (_I, [_P, C.POINTER(CtxConvParams)]),
    'ctx_prog_add_conv_tc': (_I, [_P, C.POINTER(CtxConvParams)]),
    'ctx_prog_add_pool': (_I, [_P, C.POINTER(CtxPoolParams)]),
    'ctx_prog_add_nchw_to_nhwc': (_I, [_P, _P, _P, _I, _I, _I, _I, _I]),
    'ctx_prog_add_attention': (_I, [_P, C.POINTER(CtxAttnParams)]),
    'ctx_prog_add_softmax': (_I, [_P, _P, _P, _LL, _I]),
    'ctx_prog_set_lane': (_I, [_P, _I, C.c_uint]),
    'ctx_prog_autotune': (_I, [_P, _P, _I]),
    'ctx_prog_conv_config': (_I, [_P, _I, C.POINTER(_I)]),
    'ctx_prog_num_ops': (_I, [_P]),
    'ctx_prog_run': (_I, [_P, _P]),
    'ctx_prog_instantiate_graph': (_I, [_P, _P]),
    'ctx_prog_run_range': (_I, [_P, _I, _I, _P]),
    'ctx_prog_run_range_repeat': (_I, [_P, _I, _I, _I, _P]),
    'ctx_prog_destroy': (None, [_P]),
    'ctx_match_encode': (_I, [_P, _P, _I, _P, _I, _I, _F, _F, _F, _P, _P, _P, _P, _P, _P]),
    'ctx_rank_workspace_bytes': (_SZ, [_I, _I]),
    'ctx_hard_negative_rank': (_I, [_P, _I, _I, _P, _P, _SZ, _P]),
    'ctx_loss_mining': (_I, [_P, _P, _P, _I, _I, _P, _P, _P]),
    'ctx_loss_forward_backward': (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    'ctx_point_form': (_I, [_P, _I, _P, _P]),
    'ctx_jaccard': (_I, [_P, _I, _P, _I, _P, _P]),
    'ctx_encode': (_I, [_P, _P, _I, _F, _F, _P, _P]),
    'ctx_prototype_accumulate': (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P]),
    'ctx_prototype_finalize': (_I, [_P, _P, _I, _I, _I, _P, _P]),
}

_lib = None


def lib():
    """The loaded library.  Raises if it has not been built — there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CtxError('%s is missing: build it with `python -m context_transformer_b200.build` '
                           '(there is no CPU / PyTorch fallback for the hot path)' % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)          # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(status, what=''):
    if status != 0:
        msg = lib().ctx_last_error().decode(errors='replace')
        raise CtxError('%s: %s (%s)' % (what or 'libctx_b200', msg, STATUS.get(status, status)))


def launch_count():
    return int(lib().ctx_launch_count())


def current_stream_ptr(device=None):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t, name):
    if not t.is_cuda:
        raise CtxError('%s must be a CUDA tensor: the detection hot path runs only on the GPU '
                       '(got device %s)' % (name, t.device))
    return t


def dtype_code(dt):
    import torch
    return {torch.float32: CTX_F32, torch.bfloat16: CTX_BF16, torch.float16: CTX_F16}[dt]

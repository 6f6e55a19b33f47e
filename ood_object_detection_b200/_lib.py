"""ctypes binding of libodk.so (include/odk.h).

The CUDA library is the product: there is no CPU or torch fallback.  If the shared object is
missing or a CUDA device is absent where one is needed, the shims raise immediately.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libodk.so')
_LIB = None

c_void_p, c_int, c_int64, c_float, c_double, c_size_t = (
    ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_double, ctypes.c_size_t)


MAILBOX_MAX_WORLD = 32


class Exchange(ctypes.Structure):
    """odk_exchange (include/odk.h): fused peer-mailbox exchange of the loss partial sums."""
    _fields_ = [('mailboxes', c_void_p * MAILBOX_MAX_WORLD), ('world', ctypes.c_int32), ('rank', ctypes.c_int32),
                ('num_pos_plus_1', c_void_p), ('global_out3', c_void_p), ('status', c_void_p),
                ('normalized', ctypes.c_int32), ('timeout_ms', ctypes.c_uint32)]


class LossParams(ctypes.Structure):
    _fields_ = [('alpha', c_float), ('gamma', c_float), ('delta', c_float), ('box_loss_weight', c_float),
                ('label_smoothing', c_float), ('legacy_focal', ctypes.c_int32), ('match_is_key64', ctypes.c_int32),
                ('clear_keys', ctypes.c_int32), ('layout', ctypes.c_int32), ('exchange', ctypes.POINTER(Exchange))]


class DetectParams(ctypes.Structure):
    _fields_ = [('max_det', ctypes.c_int32), ('soft_nms', ctypes.c_int32), ('score_min', c_float),
                ('nms_iou', c_double), ('soft_sigma', c_float), ('soft_iou', c_float), ('soft_score_thr', c_float),
                ('pipeline', ctypes.c_int32)]


# every symbol include/odk.h declares: name -> (restype, argtypes)
_P = c_void_p
SIGNATURES = {
    'odk_version': (c_int, []),
    'odk_last_error': (ctypes.c_char_p, []),
    'odk_planar_stride': (c_int64, [c_int64]),
    'odk_assign_workspace_bytes': (c_size_t, [c_int, c_int]),
    'odk_assign': (c_int, [_P, _P, _P, _P, c_int, c_int, _P, c_int, c_int, c_float, c_int, _P, _P, _P, c_size_t, _P]),
    'odk_assign_grid_workspace_bytes': (c_size_t, [c_int, c_int64]),
    'odk_anchor_table': (c_int, [_P, _P, c_int, _P, c_int, c_int, _P, _P]),
    'odk_assign_grid': (c_int, [_P, _P, _P, c_int, _P, _P, _P, c_int, c_int, _P, c_int, c_int, c_float, c_int, _P, _P, _P, c_int, _P, c_size_t,
                                _P]),
    'odk_keys_to_match': (c_int, [_P, c_int, c_int64, _P, _P]),
    'odk_iou_matrix': (c_int, [_P, c_int, _P, c_int, _P, _P]),
    'odk_targets': (c_int, [_P, _P, _P, c_int, c_int, _P, c_int, c_int, _P, _P, _P, _P]),
    'odk_loss_workspace_bytes': (c_size_t, []),
    'odk_loss': (c_int, [_P, _P, c_int, c_int, _P, c_int, c_int, _P, _P, _P, _P, c_int, _P, _P, _P,
                         ctypes.POINTER(LossParams), _P, _P, _P, _P, c_size_t, _P]),
    'odk_scale_inplace': (c_int, [_P, c_int64, _P, _P]),
    'odk_scale_inplace_multi': (c_int, [_P, _P, c_int, _P, _P]),
    'odk_mailbox_bytes': (c_size_t, [c_int]),
    'odk_partials_publish': (c_int, [_P, _P, c_int, c_int, _P]),
    'odk_partials_collect': (c_int, [_P, c_int, _P, _P, ctypes.c_uint32, _P]),
    'odk_topk_workspace_bytes': (c_size_t, [c_int, c_int, _P, c_int, c_int, c_int]),
    'odk_topk': (c_int, [_P, _P, c_int, c_int, _P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, c_size_t, _P]),
    'odk_detect': (c_int, [_P, _P, _P, _P, c_int, c_int, _P, c_int64, _P, _P, ctypes.POINTER(DetectParams), _P, _P, _P,
                           _P]),
    'odk_postprocess_workspace_bytes': (c_size_t, [c_int, c_int, _P, c_int, c_int, c_int]),
    'odk_postprocess_flags_offset': (c_size_t, [c_int, c_int, _P, c_int, c_int, c_int]),
    'odk_postprocess_timeline_offset': (c_size_t, [c_int, c_int, _P, c_int, c_int, c_int]),
    'odk_postprocess': (c_int, [_P, _P, c_int, c_int, _P, c_int, c_int, c_int, c_int, _P, _P, _P, ctypes.POINTER(DetectParams), c_float,
                                _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    'odk_soft_nms': (c_int, [_P, _P, c_int, c_int, c_float, c_float, c_float, c_int, _P, _P, _P, _P]),
    'odk_nms_workspace_bytes': (c_size_t, [c_int]),
    'odk_nms': (c_int, [_P, _P, c_int, c_double, _P, _P, _P, c_size_t, _P]),
    'odk_match_detections': (c_int, [_P, _P, c_int, c_int, _P, _P, _P, _P, c_int, c_int, c_int, c_double, c_double, c_int, _P, _P, _P]),
    'odk_ood': (c_int, [_P, c_int, c_int, _P, c_int, c_int, c_int, _P, c_int, c_float, _P, _P, _P]),
}


def lib():
    """Load libodk.so (once).  Fails loudly: the CUDA library is the only implementation."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                f'(or `make -C ood_object_detection_b200/csrc`). There is no CPU fallback.')
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the library does not export the symbol
            fn.restype = res
            fn.argtypes = args
        if handle.odk_version() != 3:
            raise RuntimeError('libodk.so ABI version mismatch')
        _LIB = handle
    return _LIB


def check(rc):
    if rc != 0:
        msg = lib().odk_last_error()
        raise RuntimeError(f'libodk error {rc}: {msg.decode() if msg else "?"}')


def require_cuda(t, what):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f'{what} must be a CUDA tensor: the dense per-anchor path only exists as sm_100a '
                           f'kernels (no CPU fallback)')
    return t


def ptr(t):
    return None if t is None else c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def int_array(values):
    return (ctypes.c_int32 * len(values))(*[int(v) for v in values])


def ptr_array(tensors):
    return (c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def prep_levels(outputs, num_levels, what='head output'):
    """Head outputs as the kernels read them, IN PLACE: fp32 CUDA tensors that are either contiguous NCHW or
    channels_last ([B, H, W, C] in memory -- what a channels_last / AMP head writes, efficientdet.py:405-414).
    Returns (tensors, mask) with bit l of mask set for a channels_last level; only other dtypes / stridings are
    copied."""
    outs, mask = [], 0
    for l, t in enumerate(outputs[:num_levels]):
        require_cuda(t, what)
        if t.dtype != torch.float32:
            t = t.float()
        if not t.is_contiguous():
            if t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last):
                mask |= 1 << l
            else:
                t = t.contiguous()
        outs.append(t)
    return outs, mask

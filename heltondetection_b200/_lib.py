"""ctypes binding of libhd_b200.so (the C ABI in include/hd_b200.h).

There is deliberately no CPU fallback: if the CUDA library is missing or a tensor
is not on a CUDA device the call raises.
"""
import ctypes as C
import functools
import os
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libhd_b200.so")

HD_MAX_LEVELS = 8
HD_MAX_ANCHORS = 8
FLAG_CONF_GE = 1
FLAG_DENSE_READ = 2
FLAG_MULTI_LABEL = 256
NMS_AGNOSTIC, NMS_CLASS_EXACT, NMS_CLASS_OFFSET = 0, 1, 2
LAYOUT_NCHW, LAYOUT_NHWC = 0, 1
RPN_SOFTMAX, RPN_CLAMP_DWH, RPN_KEY_LOGIT, RPN_EXACT_MATH = 1, 2, 4, 8
ROIHEAD_MUL_STD, ROIHEAD_CLAMP_DWH, ROIHEAD_LABEL_MINUS1 = 16, 32, 64
BOX_XYWH = 1
WBF_AVG, WBF_MAX, WBF_BOX_AND_MODEL_AVG, WBF_ABSENT_MODEL_AWARE_AVG, WBF_RESCALE_SUM_WEIGHTS = 0, 1, 2, 3, 256


class YoloLevel(C.Structure):
    _fields_ = [("data", C.c_void_p), ("H", C.c_int32), ("W", C.c_int32), ("stride", C.c_float),
                ("anchor_wh", C.c_float * (2 * HD_MAX_ANCHORS))]


class RpnLevel(C.Structure):
    _fields_ = [("objectness", C.c_void_p), ("deltas", C.c_void_p), ("H", C.c_int32), ("W", C.c_int32),
                ("stride", C.c_float), ("anchor_base", C.c_float * (4 * HD_MAX_ANCHORS))]


HD_MAX_REPLICAS = 16


class Replicas(C.Structure):
    _fields_ = [("n", C.c_int32), ("det", C.c_void_p * HD_MAX_REPLICAS), ("count", C.c_void_p * HD_MAX_REPLICAS)]


class RoiLevel(C.Structure):
    _fields_ = [("data", C.c_void_p), ("H", C.c_int32), ("W", C.c_int32), ("spatial_scale", C.c_float)]


_vp, _i, _i64, _f, _d, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_size_t

# name -> (restype, argtypes); every symbol include/hd_b200.h declares appears here
SIGNATURES = {
    "hd_version": (_i, []),
    "hd_last_error": (C.c_char_p, []),
    "hd_debug_phases": (_i, [_i, C.POINTER(C.c_longlong)]),
    "hd_debug_launch_count": (C.c_ulonglong, []),
    "hd_debug_roi_profile": (_i, [_i, C.POINTER(C.c_ulonglong)]),
    "hd_debug_roi_align_sliced": (_i, [C.POINTER(RoiLevel), _i, _i, _vp, _vp, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp]),
    "hd_yolo_decode": (_i, [C.POINTER(YoloLevel), _i, _i, _i, _i, _vp, _vp]),
    "hd_yolo_decode_filter": (_i, [C.POINTER(YoloLevel), _i, _i, _i, _i, _d, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "hd_yolo_postprocess_workspace_size": (_sz, [_i, _i]),
    "hd_yolo_postprocess": (_i, [C.POINTER(YoloLevel), _i, _i, _i, _i, _d, _d, _i, _i, _f, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "hd_yolo_filter_pred": (_i, [_vp, _i, _i, _i, _d, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "hd_sort_nms_workspace_size": (_sz, [_i, _i]),
    "hd_nms_set_mode": (_i, [_i]),
    "hd_sort_nms_batched": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _d, _i, _f, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "hd_sort_nms_batched_replicated": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _d, _i, _f, _i, _i, _vp, _vp, _vp, C.POINTER(Replicas), _vp, _sz, _vp]),
    "hd_yolo_postprocess_replicated": (_i, [C.POINTER(YoloLevel), _i, _i, _i, _i, _d, _d, _i, _i, _f, _i, _i, _vp, _vp, _vp, C.POINTER(Replicas), _vp, _sz, _vp]),
    "hd_box_iou": (_i, [_vp, _i64, _vp, _i64, _vp, _vp]),
    "hd_rpn_num_anchors": (_i, [C.POINTER(RpnLevel), _i, _i]),
    "hd_rpn_decode": (_i, [C.POINTER(RpnLevel), _i, _i, _i, _i, _f, _f, _f, _f, _vp, _vp, _vp, _vp]),
    "hd_rpn_select_nms_workspace_size": (_sz, [_i, _i, _i]),
    "hd_rpn_select_nms": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _d, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "hd_rpn_set_mode": (_i, [_i]),
    "hd_rpn_cluster_capacity": (_i, [_i]),
    "hd_rpn_set_cluster_size": (_i, [_i]),
    "hd_rpn_proposals_workspace_size": (_sz, [_i, _i, _i]),
    "hd_rpn_proposals": (_i, [C.POINTER(RpnLevel), _i, _i, _i, _i, _f, _f, _f, _f, _i, _i, _d, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "hd_rpn_merge_levels": (_i, [_vp, _vp, C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(C.c_int32), _i, _i, _i, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hd_rpn_select_nms_strided": (_i, [_vp, _vp, _vp, _i, _i, _i64, _i, _i, _d, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "hd_rpn_finish_levels": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "hd_roi_head_decode_filter": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, C.POINTER(C.c_float), _i, _f, _f, _f, _d, _f, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "hd_roi_head_postprocess_workspace_size": (_sz, [_i, _i, _i]),
    "hd_roi_head_postprocess": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, C.POINTER(C.c_float), _i, _f, _f, _f, _d, _f, _d, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "hd_scale_detections": (_i, [_vp, _vp, _i, _i, _vp, _i, _vp, _vp]),
    "hd_box_encode": (_i, [_vp, _i, _i, _vp, _vp, _i, _i, C.POINTER(C.c_float), _vp, _vp]),
    "hd_match_workspace_size": (_sz, [_i, _i, _i]),
    "hd_match": (_i, [_vp, _vp, _i, _i, _vp, _i, _i, _d, _d, _i, _vp, _vp, _vp, _sz, _vp]),
    "hd_wbf_workspace_size": (_sz, [_i, _i, _i, _i]),
    "hd_wbf": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, C.POINTER(C.c_double), _d, _d, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "hd_tta_map_back": (_i, [_vp, _vp, _i, _i, _f, _i, _f, _f, _f, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "hd_roi_align": (_i, [C.POINTER(RoiLevel), _i, _i, _i, _vp, _vp, _i64, _i, _i, _i, _i, _vp, _vp]),
    "hd_roi_align_workspace_size": (_sz, [C.POINTER(RoiLevel), _i, _i, _i, _i64, _i, _i, _i]),
    "hd_roi_align_ws": (_i, [C.POINTER(RoiLevel), _i, _i, _i, _i, _vp, _vp, _i64, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "hd_roi_pool": (_i, [C.POINTER(RoiLevel), _i, _i, _i, _vp, _vp, _i64, _i, _i, _vp, _vp, _vp]),
    "hd_roi_set_mode": (_i, [_i]),
    "hd_roi_pool_backward": (_i, [_vp, _vp, _vp, _i64, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "hd_roi_align_backward": (_i, [_vp, _vp, _vp, _i64, C.POINTER(RoiLevel), _i, _i, _i, _i, _i, _i, _i, _vp]),
    "hd_nchw_to_nhwc": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "hd_roi_level_map": (_i, [_vp, _i, _i, _i64, _i, _i, _i, _f, _f, _f, _vp, _vp, _vp]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C heltondetection_b200/csrc`). There is no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError(lib().hd_last_error().decode())


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _first_cuda_device(objs):
    for o in objs:
        if isinstance(o, torch.Tensor):
            if o.is_cuda:
                return o.device
        elif isinstance(o, (list, tuple)):
            d = _first_cuda_device(o)
            if d is not None:
                return d
    return None


def on_device(fn):
    """Run `fn` with the device of its first CUDA tensor argument current (as torchvision's ops do with a device guard):
    the C ABI launches on the current device, and the stream handed to it must belong to that device."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = _first_cuda_device(args) or _first_cuda_device(tuple(kwargs.values()))
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapper


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("heltondetection_b200 ops are CUDA-only (sm_100a); got a tensor on " + str(t.device))


def f32c(t):
    """float32 contiguous view/copy (inputs are borrowed, never mutated)."""
    if t.dtype != torch.float32:
        raise NotImplementedError(f"heltondetection_b200 kernels are fp32-only, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()

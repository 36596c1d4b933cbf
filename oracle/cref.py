"""ctypes view of oracle/c/libhd_oracle.so (plain-C restatement; built by __graft_entry__.build() / make -C oracle/c)."""
import ctypes as C
import os
import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c", "libhd_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            import subprocess
            subprocess.check_call(["make", "-C", os.path.dirname(_PATH)])
        _lib = C.CDLL(_PATH)
        _lib.hdo_nms.restype = C.c_int64
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, np.float32)
    return a, a.ctypes.data_as(C.c_void_p)


def nms(boxes, scores, thr, max_keep=-1):
    b, bp = _f(boxes)
    s, sp = _f(scores)
    keep = np.empty(len(s), np.int64)
    k = lib().hdo_nms(bp, sp, C.c_int64(len(s)), C.c_double(thr), C.c_int64(max_keep), keep.ctypes.data_as(C.c_void_p))
    return keep[:k]


def box_iou(b1, b2):
    a, ap = _f(b1)
    b, bp = _f(b2)
    out = np.empty((len(a), len(b)), np.float32)
    lib().hdo_box_iou(ap, C.c_int64(len(a)), bp, C.c_int64(len(b)), out.ctypes.data_as(C.c_void_p))
    return out


def roi_align(x, rois, output_size, scale, sampling_ratio, aligned):
    x, xp = _f(x)
    r, rp = _f(rois)
    PH, PW = (output_size, output_size) if isinstance(output_size, int) else output_size
    out = np.empty((len(r), x.shape[1], PH, PW), np.float32)
    lib().hdo_roi_align(xp, x.shape[1], x.shape[2], x.shape[3], rp, C.c_int64(len(r)), C.c_float(scale), PH, PW, int(sampling_ratio),
                        int(bool(aligned)), out.ctypes.data_as(C.c_void_p))
    return out


def roi_pool(x, rois, output_size, scale):
    x, xp = _f(x)
    r, rp = _f(rois)
    PH, PW = (output_size, output_size) if isinstance(output_size, int) else output_size
    out = np.empty((len(r), x.shape[1], PH, PW), np.float32)
    lib().hdo_roi_pool(xp, x.shape[1], x.shape[2], x.shape[3], rp, C.c_int64(len(r)), C.c_float(scale), PH, PW, out.ctypes.data_as(C.c_void_p))
    return out


def match(gt, pred, high, low, allow_low):
    g, gp = _f(gt)
    p, pp = _f(pred)
    out = np.empty(len(p), np.int64)
    lib().hdo_match(gp, C.c_int64(len(g)), pp, C.c_int64(len(p)), C.c_double(high), C.c_double(low), int(bool(allow_low)),
                    out.ctypes.data_as(C.c_void_p))
    return out


def scale_coords(det, pad_x, pad_y, gain, w0, h0, xywh=False):
    d, dp = _f(det)
    out = np.empty_like(d)
    lib().hdo_scale_coords(dp, C.c_int64(len(d)), C.c_float(pad_x), C.c_float(pad_y), C.c_float(gain), C.c_float(w0), C.c_float(h0), int(bool(xywh)),
                           out.ctypes.data_as(C.c_void_p))
    return out

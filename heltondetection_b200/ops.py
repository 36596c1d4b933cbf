"""torchvision-compatible box ops on the hd_b200 kernels.

``nms`` / ``batched_nms`` / ``box_iou`` keep the torchvision signatures, argument meaning and
error messages (torchvision/ops/boxes.py:20-120, 308-370).  They are also registered as the
``hd_b200::nms`` / ``hd_b200::box_iou`` PyTorch custom ops with the torchvision schemas.
"""
import torch
from . import _lib


def set_nms_mode(mode):
    """large-segment NMS kernel choice: 0 auto (thread-block clusters for small batches), 1 one CTA per image, 2 clusters whenever
    possible, 3 always the light 256-thread kernel, 4 always the 1024-thread kernel, 5 as 1 with the bitonic network instead of the
    bucket sort; returns the previous mode."""
    return _lib.lib().hd_nms_set_mode(int(mode))


def _check_nms_args(boxes, scores):
    if boxes.dim() != 2:
        raise RuntimeError(f"boxes should be a 2d tensor, got {boxes.dim()}D")
    if boxes.size(1) != 4:
        raise RuntimeError(f"boxes should have 4 elements in dimension 1, got {boxes.size(1)}")
    if scores.dim() != 1:
        raise RuntimeError(f"scores should be a 1d tensor, got {scores.dim()}D")
    if boxes.size(0) != scores.size(0):
        raise RuntimeError("boxes and scores should have same number of elements in dimension 0, "
                           f"got {boxes.size(0)} and {scores.size(0)}")
    if boxes.dtype != scores.dtype:
        raise RuntimeError("dets should have the same type as scores")
    if boxes.dtype != torch.float32:
        raise NotImplementedError(f'"nms_kernel" not implemented for \'{str(boxes.dtype).split(".")[-1].capitalize()}\' (hd_b200 is fp32-only)')


@_lib.on_device
def _nms_single(boxes, scores, iou_threshold, cls=None, mode=_lib.NMS_AGNOSTIC, offset=0.0, max_det=None, max_nms=0):
    """One image through hd_sort_nms_batched -> int64 keep indices in score order."""
    _lib.require_cuda(boxes, scores, cls)
    _check_nms_args(boxes, scores)
    n = boxes.shape[0]
    if n == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    boxes, scores = _lib.f32c(boxes), _lib.f32c(scores)
    if cls is not None:
        cls = cls.to(torch.int32).contiguous()
    md = n if max_det is None else int(max_det)
    L = _lib.lib()
    ws_bytes = L.hd_sort_nms_workspace_size(1, n)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=boxes.device)
    idx = torch.empty((md,), dtype=torch.int64, device=boxes.device)
    cnt = torch.zeros((1,), dtype=torch.int32, device=boxes.device)
    _lib.check(L.hd_sort_nms_batched(_lib.ptr(boxes), _lib.ptr(scores), _lib.ptr(cls), None, None, n, 1, n,
                                     float(iou_threshold), mode, float(offset), int(max_nms), md, None, _lib.ptr(idx),
                                     _lib.ptr(cnt), _lib.ptr(ws), ws_bytes, _lib.stream()))
    return idx[: int(cnt.item())]  # data-dependent size: the op's one host sync, as in torchvision


def nms(boxes, scores, iou_threshold):
    """torchvision.ops.nms (boxes.py:20-48)."""
    return torch.ops.hd_b200.nms(boxes, scores, float(iou_threshold))


def batched_nms(boxes, scores, idxs, iou_threshold):
    """torchvision.ops.batched_nms (boxes.py:51-120) with exact class masking
    (= _batched_nms_vanilla for every input size; no coordinate-offset rounding)."""
    _lib.require_cuda(boxes, scores, idxs)
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    return _nms_single(boxes, scores, iou_threshold, idxs, _lib.NMS_CLASS_EXACT)


def box_iou(boxes1, boxes2):
    """torchvision.ops.box_iou (boxes.py:308-370)."""
    return torch.ops.hd_b200.box_iou(boxes1, boxes2)


@_lib.on_device
def _box_iou_impl(boxes1, boxes2):
    _lib.require_cuda(boxes1, boxes2)
    if boxes1.dim() != 2 or boxes1.size(1) != 4 or boxes2.dim() != 2 or boxes2.size(1) != 4:
        raise RuntimeError("box_iou expects Tensor[N, 4] and Tensor[M, 4]")
    b1, b2 = _lib.f32c(boxes1), _lib.f32c(boxes2)
    out = torch.empty((b1.shape[0], b2.shape[0]), dtype=torch.float32, device=b1.device)
    _lib.check(_lib.lib().hd_box_iou(_lib.ptr(b1), b1.shape[0], _lib.ptr(b2), b2.shape[0], _lib.ptr(out), _lib.stream()))
    return out


# ----------------------------------------------------------------------------- custom-op registration
_LIB = torch.library.Library("hd_b200", "DEF")
_LIB.define("nms(Tensor dets, Tensor scores, float iou_threshold) -> Tensor")
_LIB.define("box_iou(Tensor boxes1, Tensor boxes2) -> Tensor")


def _cpu_refuse(*a, **k):
    raise RuntimeError("hd_b200 ops are CUDA-only (sm_100a); there is no CPU fallback")


_LIB.impl("nms", lambda dets, scores, thr: _nms_single(dets, scores, thr), "CUDA")
_LIB.impl("box_iou", _box_iou_impl, "CUDA")
_LIB.impl("nms", _cpu_refuse, "CPU")
_LIB.impl("box_iou", _cpu_refuse, "CPU")

# RoI ops re-exported beside the box ops, as in torchvision.ops
from .roi import (roi_align, roi_pool, roi_pool_with_argmax, RoIAlign, RoIPool, level_map,  # noqa: E402,F401
                  multilevel_roi_align, convert_boxes_to_roi_format)

_LIB.define("roi_align(Tensor input, Tensor rois, float spatial_scale, SymInt pooled_height, SymInt pooled_width, "
            "int sampling_ratio, bool aligned) -> Tensor")
_LIB.define("roi_pool(Tensor input, Tensor rois, float spatial_scale, SymInt pooled_height, SymInt pooled_width) -> (Tensor, Tensor)")
_LIB.impl("roi_align", lambda input, rois, s, ph, pw, sr, al: roi_align(input, rois, (ph, pw), s, sr, al), "CUDA")
_LIB.impl("roi_pool", lambda input, rois, s, ph, pw: roi_pool_with_argmax(input, rois, (ph, pw), s), "CUDA")
_LIB.impl("roi_align", _cpu_refuse, "CPU")
_LIB.impl("roi_pool", _cpu_refuse, "CPU")


# ----------------------------------------------------------------------------- fake (meta) implementations
# Shapes/dtypes without data, so that the four ops trace under FakeTensorMode / torch.compile / torch.export exactly like the
# torchvision ops they replace (torchvision registers the same abstract impls in torchvision/_meta_registrations.py).
@torch.library.register_fake("hd_b200::nms")
def _nms_fake(dets, scores, iou_threshold):
    torch._check(dets.dim() == 2, lambda: f"boxes should be a 2d tensor, got {dets.dim()}D")
    torch._check(dets.size(1) == 4, lambda: f"boxes should have 4 elements in dimension 1, got {dets.size(1)}")
    torch._check(scores.dim() == 1, lambda: f"scores should be a 1d tensor, got {scores.dim()}D")
    torch._check(dets.size(0) == scores.size(0),
                 lambda: f"boxes and scores should have same number of elements in dimension 0, got {dets.size(0)} and {scores.size(0)}")
    n = torch.library.get_ctx().new_dynamic_size()      # data-dependent number of kept boxes
    return dets.new_empty((n,), dtype=torch.int64)


@torch.library.register_fake("hd_b200::box_iou")
def _box_iou_fake(boxes1, boxes2):
    torch._check(boxes1.dim() == 2 and boxes1.size(1) == 4 and boxes2.dim() == 2 and boxes2.size(1) == 4,
                 lambda: "box_iou expects Tensor[N, 4] and Tensor[M, 4]")
    return boxes1.new_empty((boxes1.size(0), boxes2.size(0)))


@torch.library.register_fake("hd_b200::roi_align")
def _roi_align_fake(input, rois, spatial_scale, pooled_height, pooled_width, sampling_ratio, aligned):
    torch._check(rois.dim() == 2 and rois.size(1) == 5, lambda: "rois must have shape as Tensor[K, 5]")
    return input.new_empty((rois.size(0), input.size(1), pooled_height, pooled_width))


@torch.library.register_fake("hd_b200::roi_pool")
def _roi_pool_fake(input, rois, spatial_scale, pooled_height, pooled_width):
    torch._check(rois.dim() == 2 and rois.size(1) == 5, lambda: "rois must have shape as Tensor[K, 5]")
    shape = (rois.size(0), input.size(1), pooled_height, pooled_width)
    return input.new_empty(shape), input.new_empty(shape, dtype=torch.int32)

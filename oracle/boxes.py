"""IoU / NMS oracle.

Primary functions call the torchvision CPU ops the reference calls
(torchvision/ops/boxes.py:20-48 nms, :51-120 batched_nms, :308-370 box_iou).
``*_restated`` are independent numpy restatements used to show the semantics
are understood; tests pin them to the torchvision ops.
"""
import numpy as np
import torch
import torchvision


def box_iou(boxes1: torch.Tensor, boxes2: torch.Tensor) -> torch.Tensor:
    """torchvision/ops/boxes.py:308-370 (area :295-297, inter/union :321-339)."""
    return torchvision.ops.box_iou(boxes1.cpu(), boxes2.cpu())


def box_iou_restated(b1: np.ndarray, b2: np.ndarray) -> np.ndarray:
    """fp32, op order of boxes.py:321-339,369: inter / (area1 + area2 - inter), no eps."""
    b1 = np.asarray(b1, np.float32)
    b2 = np.asarray(b2, np.float32)
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    lt = np.maximum(b1[:, None, :2], b2[None, :, :2])
    rb = np.minimum(b1[:, None, 2:], b2[None, :, 2:])
    wh = np.clip(rb - lt, 0, None)
    inter = wh[..., 0] * wh[..., 1]
    union = (a1[:, None] + a2[None, :]) - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        return (inter / union).astype(np.float32)


def nms(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """torchvision.ops.nms CPU kernel (boxes.py:20-48)."""
    return torchvision.ops.nms(boxes.cpu(), scores.cpu(), float(iou_threshold))


def nms_restated(boxes: np.ndarray, scores: np.ndarray, iou_threshold: float, max_keep: int = -1) -> np.ndarray:
    """Greedy NMS, fp32: stable descending sort (ties -> lower index, NaN score first),
    suppress iff IoU > thr (strict).  The C++ CPU kernel compares the fp32 IoU
    against the *double* threshold (float promoted), so the compare is done in fp64."""
    boxes = np.asarray(boxes, np.float32)
    scores = np.asarray(scores, np.float32)
    n = boxes.shape[0]
    # torch.sort(descending=True) puts NaN first and is stable on CPU
    key = np.where(np.isnan(scores), np.float32(np.inf), scores)
    isn = np.isnan(scores)
    order = np.lexsort((np.arange(n), -key.astype(np.float64), ~isn))
    x1, y1, x2, y2 = (boxes[:, i] for i in range(4))
    areas = (x2 - x1) * (y2 - y1)
    thr = float(iou_threshold)
    suppressed = np.zeros(n, bool)
    keep = []
    for _i in range(n):
        i = order[_i]
        if suppressed[i]:
            continue
        keep.append(i)
        if 0 <= max_keep == len(keep):
            break
        rest = order[_i + 1:]
        xx1 = np.maximum(x1[i], x1[rest])
        yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest])
        yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(np.float32(0), xx2 - xx1)
        h = np.maximum(np.float32(0), yy2 - yy1)
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / ((areas[i] + areas[rest]) - inter)
        suppressed[rest[ovr.astype(np.float64) > thr]] = True
    return np.asarray(keep, np.int64)


def batched_nms(boxes, scores, idxs, iou_threshold):
    """Exact per-class semantics: torchvision _batched_nms_vanilla (boxes.py:103-120).
    The coordinate-trick variant (:85-100) is torchvision.ops.batched_nms itself for
    small inputs; both are exposed so tests can compare."""
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64)
    return torchvision.ops.boxes._batched_nms_vanilla(boxes.cpu(), scores.cpu(), idxs.cpu(), float(iou_threshold))


def batched_nms_offset(boxes, scores, idxs, iou_threshold, offset_scale):
    """Coordinate-offset class separation with a fixed scale (ultralytics max_wh, A.2):
    nms(boxes + idxs*offset_scale) with the add done in fp32."""
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64)
    off = idxs.to(boxes.dtype).cpu() * boxes.new_tensor(offset_scale).cpu()
    return torchvision.ops.nms(boxes.cpu() + off[:, None], scores.cpu(), float(iou_threshold))

"""IoU-based label assignment: anchor / proposal <-> ground-truth matching.

``Matcher`` keeps torchvision ``det_utils.Matcher`` semantics (models/detection/_utils.py:318-400) but takes the boxes
instead of a materialised IoU matrix; ``anchor_labels`` reads the same matching as the lineage AnchorTargetCreator
label rule (1 positive, 0 negative, -1 ignored; bubbliiiing frcnn utils_fit), without its random subsampling."""
import ctypes as C
import torch
from . import _lib

BELOW_LOW_THRESHOLD, BETWEEN_THRESHOLDS = -1, -2


class Matcher:
    def __init__(self, high_threshold, low_threshold, allow_low_quality_matches=False):
        if low_threshold > high_threshold:
            raise AssertionError("low_threshold should be <= high_threshold")
        self.high, self.low, self.allow = float(high_threshold), float(low_threshold), bool(allow_low_quality_matches)

    @_lib.on_device
    def __call__(self, gt_boxes, boxes, gt_count=None, return_iou=False):
        """gt_boxes [G,4] | [B,Gmax,4] (+ gt_count [B]), boxes [N,4] (shared anchors) | [B,N,4] -> matches [N] | [B,N] int64"""
        _lib.require_cuda(gt_boxes, boxes, gt_count)
        single = gt_boxes.dim() == 2
        gt = _lib.f32c(gt_boxes[None] if single else gt_boxes)
        pr = _lib.f32c(boxes)
        B, G = gt.shape[0], gt.shape[1]
        if G == 0:
            raise ValueError("No ground-truth boxes available for one of the images during training")
        per_image = pr.dim() == 3
        if per_image and pr.shape[0] != B:
            raise RuntimeError(f"boxes batch {pr.shape[0]} != gt batch {B}")
        N = pr.shape[-2]
        L = _lib.lib()
        ws_bytes = L.hd_match_workspace_size(B, G, N)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=gt.device)
        matches = torch.empty((B, N), dtype=torch.int64, device=gt.device)
        iou = torch.empty((B, N), dtype=torch.float32, device=gt.device) if return_iou else None
        if gt_count is not None:
            gt_count = gt_count.to(torch.int32).contiguous()
        _lib.check(L.hd_match(_lib.ptr(gt), _lib.ptr(gt_count), B, G, _lib.ptr(pr), 1 if per_image else 0, N, self.high, self.low,
                              1 if self.allow else 0, _lib.ptr(matches), _lib.ptr(iou), _lib.ptr(ws), ws_bytes, _lib.stream()))
        if single:
            matches, iou = matches[0], (iou[0] if iou is not None else None)
        return (matches, iou) if return_iou else matches


def anchor_labels(gt_boxes, anchors, pos_iou_thresh=0.7, neg_iou_thresh=0.3, gt_count=None):
    """lineage AnchorTargetCreator._create_label without the random subsampling -> (argmax_gt [.., N], label [.., N] in {1, 0, -1})"""
    m = Matcher(pos_iou_thresh, neg_iou_thresh, True)(gt_boxes, anchors, gt_count)
    label = torch.where(m >= 0, torch.ones_like(m), torch.where(m == BELOW_LOW_THRESHOLD, torch.zeros_like(m), -torch.ones_like(m)))
    return m.clamp(min=0), label


@_lib.on_device
def encode_boxes(gt_boxes, boxes, matches=None, weights=(1.0, 1.0, 1.0, 1.0)):
    """torchvision BoxCoder.encode_single / lineage bbox2loc: regression targets of `boxes` towards their matched ground truth.
    gt_boxes [G,4] | [B,G,4]; boxes [N,4] | [B,N,4]; matches [N] | [B,N] from Matcher (negative codes are read as GT 0, as
    torchvision's clamp(min=0) does); matches=None encodes row i against row i."""
    _lib.require_cuda(gt_boxes, boxes, matches)
    single = gt_boxes.dim() == 2
    gt = _lib.f32c(gt_boxes[None] if single else gt_boxes)
    pr = _lib.f32c(boxes)
    B, G = gt.shape[0], gt.shape[1]
    per_image = pr.dim() == 3
    N = pr.shape[-2]
    out = torch.empty((B, N, 4), dtype=torch.float32, device=gt.device)
    m = None
    if matches is not None:
        m = (matches[None] if matches.dim() == 1 else matches).to(torch.int64).contiguous()
    w = (C.c_float * 4)(*[float(x) for x in weights])
    _lib.check(_lib.lib().hd_box_encode(_lib.ptr(gt), B, G, _lib.ptr(m), _lib.ptr(pr), 1 if per_image else 0, N, w, _lib.ptr(out), _lib.stream()))
    return out[0] if single else out

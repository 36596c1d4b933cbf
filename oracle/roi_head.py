"""RoI-head output post-process oracle (reference feature README.md:8,61).

Pinned at the executable reference of the lineage: ``postprocess_detections_tv`` CALLS torchvision's own
``RoIHeads.postprocess_detections`` (models/detection/roi_heads.py:668-723, BoxCoder.decode_single
_utils.py:183-224) on CPU; ``postprocess_detections`` restates it with the lineage variants as options
(std multipliers, label-1, no clamp: bubbliiiing DecodeBox) and is checked against the former in
tests/test_oracle.py.  ``scale_coords`` / ``xyxy2xywh`` follow ultralytics general.py (letterbox inverse) and the
COCO result-json box format.
"""
import math
import torch
import torchvision
from torchvision.models.detection.roi_heads import RoIHeads

BBOX_XFORM_CLIP = math.log(1000.0 / 16)


def postprocess_detections_tv(class_logits, box_regression, proposals, image_shapes, score_thresh=0.05, nms_thresh=0.5,
                              detections_per_img=100, weights=(10.0, 10.0, 5.0, 5.0)):
    """the unmodified torchvision method (only box_coder / thresholds are read by it)"""
    rh = RoIHeads(None, None, None, 0.5, 0.5, 512, 0.25, weights, score_thresh, nms_thresh, detections_per_img)
    return rh.postprocess_detections(class_logits, box_regression, [p.float() for p in proposals], image_shapes)


def decode(rel, boxes, weights, mul_std=False, clamp=BBOX_XFORM_CLIP):
    """[R, C*4] deltas on [R,4] boxes -> [R, C, 4]"""
    w = boxes[:, 2] - boxes[:, 0]
    h = boxes[:, 3] - boxes[:, 1]
    cx = boxes[:, 0] + 0.5 * w
    cy = boxes[:, 1] + 0.5 * h
    wx, wy, ww, wh = weights
    if mul_std:
        dx, dy, dw, dh = rel[:, 0::4] * wx, rel[:, 1::4] * wy, rel[:, 2::4] * ww, rel[:, 3::4] * wh
    else:
        dx, dy, dw, dh = rel[:, 0::4] / wx, rel[:, 1::4] / wy, rel[:, 2::4] / ww, rel[:, 3::4] / wh
    if clamp is not None:
        dw, dh = torch.clamp(dw, max=clamp), torch.clamp(dh, max=clamp)
    pcx, pcy = dx * w[:, None] + cx[:, None], dy * h[:, None] + cy[:, None]
    pw, ph = torch.exp(dw) * w[:, None], torch.exp(dh) * h[:, None]
    hw, hh = 0.5 * pw, 0.5 * ph
    return torch.stack((pcx - hw, pcy - hh, pcx + hw, pcy + hh), 2)


def candidates(class_logits, box_regression, proposals_img, image_shape, score_thresh, weights, mul_std=False, clamp=BBOX_XFORM_CLIP,
               min_size=1e-2, ge=False, label_minus1=False):
    """one image -> (boxes [n,4], scores [n], labels [n], ids [n]) in flat (roi, class) order"""
    C = class_logits.shape[-1]
    boxes = decode(box_regression, proposals_img.float(), weights, mul_std, clamp)
    scores = torch.softmax(class_logits, -1)
    boxes = torchvision.ops.clip_boxes_to_image(boxes, image_shape)
    labels = torch.arange(C).view(1, -1).expand_as(scores)
    boxes, scores, labels = boxes[:, 1:].reshape(-1, 4), scores[:, 1:].reshape(-1), labels[:, 1:].reshape(-1)
    ids = torch.arange(scores.numel())
    keep = (scores >= score_thresh) if ge else (scores > score_thresh)
    if min_size is not None:
        keep &= ((boxes[:, 2] - boxes[:, 0]) >= min_size) & ((boxes[:, 3] - boxes[:, 1]) >= min_size)
    boxes, scores, labels, ids = boxes[keep], scores[keep], labels[keep], ids[keep]
    if label_minus1:
        labels = labels - 1
    return boxes, scores, labels, ids


def postprocess_detections(class_logits, box_regression, proposals, image_shapes, score_thresh=0.05, nms_thresh=0.5, detections_per_img=100,
                           weights=(10.0, 10.0, 5.0, 5.0), mul_std=False, clamp=BBOX_XFORM_CLIP, min_size=1e-2, ge=False, label_minus1=False,
                           return_ids=False):
    """restatement; returns (boxes, scores, labels[, ids]) lists like the torchvision method"""
    n = [p.shape[0] for p in proposals]
    out_b, out_s, out_l, out_i = [], [], [], []
    for lg, rg, pr, shp in zip(class_logits.split(n, 0), box_regression.split(n, 0), proposals, image_shapes):
        b, s, l, i = candidates(lg, rg, pr, shp, score_thresh, weights, mul_std, clamp, min_size, ge, label_minus1)
        keep = torchvision.ops.boxes._batched_nms_vanilla(b, s, l, nms_thresh) if b.numel() else torch.zeros((0,), dtype=torch.int64)
        keep = keep[:detections_per_img]
        out_b.append(b[keep]); out_s.append(s[keep]); out_l.append(l[keep]); out_i.append(i[keep])
    return (out_b, out_s, out_l, out_i) if return_ids else (out_b, out_s, out_l)


def scale_coords(img1_shape, coords, img0_shape):
    """ultralytics scale_coords + clip_coords: letterboxed (img1) xyxy -> original image (img0) pixels, fp32 tensor arithmetic"""
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad = (img1_shape[1] - img0_shape[1] * gain) / 2, (img1_shape[0] - img0_shape[0] * gain) / 2
    c = coords.clone().float()
    c[:, [0, 2]] -= pad[0]
    c[:, [1, 3]] -= pad[1]
    c[:, :4] /= gain
    c[:, [0, 2]] = c[:, [0, 2]].clamp(0, img0_shape[1])
    c[:, [1, 3]] = c[:, [1, 3]].clamp(0, img0_shape[0])
    return c


def xyxy2xywh_coco(boxes):
    """COCO result-json box: [x_min, y_min, width, height]"""
    out = boxes.clone()
    out[:, 2] = boxes[:, 2] - boxes[:, 0]
    out[:, 3] = boxes[:, 3] - boxes[:, 1]
    return out

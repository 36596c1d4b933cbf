"""RoI-head output post-process and output formats (reference feature README.md:8,61).

``postprocess_detections`` keeps the signature of torchvision ``RoIHeads.postprocess_detections``
(models/detection/roi_heads.py:668) -- softmax, per-class decode (weights 10,10,5,5), clip, score / size filter,
class-aware NMS, top-k -- and ``DecodeBox`` the lineage call (bubbliiiing frcnn ``DecodeBox.forward``: std
multipliers 0.1,0.1,0.2,0.2, labels c-1, results grouped by class).  ``RoIHeadPostprocessor`` is the padded,
sync-free form that consumes ``RpnProposals`` output directly.  ``scale_coords`` is the letterbox inverse
(ultralytics general.py) on padded detections; ``coco_records`` builds the result-json rows on the host.
"""
import ctypes as C
import math
import torch
from . import _lib

BBOX_XFORM_CLIP = math.log(1000.0 / 16)


class RoIHeadPostprocessor:
    """(class_logits [B*R,C], box_regression [B*R,4C], rois [B*R,5], roi_count [B]|None) -> (det [B,max_det,6], idx, count); no host sync."""

    def __init__(self, image_size, score_thresh=0.05, nms_thresh=0.5, detections_per_img=100, weights=(10.0, 10.0, 5.0, 5.0),
                 mul_std=False, clamp=BBOX_XFORM_CLIP, min_size=1e-2, ge=False, label_minus1=False):
        self.image_size = (float(image_size[0]), float(image_size[1]))
        self.score_thresh, self.nms_thresh, self.max_det = float(score_thresh), float(nms_thresh), int(detections_per_img)
        self.weights = (C.c_float * 4)(*[float(w) for w in weights])
        self.flags = ((_lib.ROIHEAD_MUL_STD if mul_std else 0) | (_lib.ROIHEAD_CLAMP_DWH if clamp is not None else 0)
                      | (_lib.ROIHEAD_LABEL_MINUS1 if label_minus1 else 0) | (_lib.FLAG_CONF_GE if ge else 0))
        self.clamp = float(clamp) if clamp is not None else 0.0
        self.min_size = float(min_size) if min_size is not None else -math.inf
        self._key = None

    def _alloc(self, B, R, nclass, dev):
        key = (B, R, nclass, dev)
        if self._key != key:
            L = _lib.lib()
            self.ws_bytes = L.hd_roi_head_postprocess_workspace_size(B, R, nclass)
            self.ws = torch.empty((self.ws_bytes,), dtype=torch.uint8, device=dev)
            self.det = torch.zeros((B, self.max_det, 6), dtype=torch.float32, device=dev)
            self.idx = torch.empty((B, self.max_det), dtype=torch.int64, device=dev)
            self.count = torch.zeros((B,), dtype=torch.int32, device=dev)
            self._key = key

    @_lib.on_device
    def __call__(self, class_logits, box_regression, rois, roi_count=None, B=None):
        _lib.require_cuda(class_logits, box_regression, rois, roi_count)
        class_logits, box_regression, rois = _lib.f32c(class_logits), _lib.f32c(box_regression), _lib.f32c(rois)
        nclass = class_logits.shape[-1]
        if box_regression.shape[-1] != 4 * nclass or box_regression.shape[0] != class_logits.shape[0] or rois.shape[0] != class_logits.shape[0]:
            raise RuntimeError(f"expected class_logits [N,{nclass}], box_regression [N,{4 * nclass}] and rois [N,5]; got "
                               f"{tuple(class_logits.shape)}, {tuple(box_regression.shape)}, {tuple(rois.shape)}")
        if B is None:
            B = int(roi_count.numel()) if roi_count is not None else 1
        N = class_logits.shape[0]
        if B <= 0 or N % B:
            raise RuntimeError(f"{N} rows cannot be split into {B} images of equal (padded) length")
        R = N // B
        if roi_count is not None:
            roi_count = roi_count.to(torch.int32).contiguous()
        self._alloc(B, R, nclass, class_logits.device)
        _lib.check(_lib.lib().hd_roi_head_postprocess(
            _lib.ptr(class_logits), _lib.ptr(box_regression), _lib.ptr(rois), _lib.ptr(roi_count), B, R, nclass, self.weights, self.flags,
            self.clamp, self.image_size[0], self.image_size[1], self.score_thresh, self.min_size, self.nms_thresh, self.max_det,
            _lib.ptr(self.det), _lib.ptr(self.idx), _lib.ptr(self.count), _lib.ptr(self.ws), self.ws_bytes, _lib.stream()))
        return self.det, self.idx, self.count


def _pad_lists(class_logits, box_regression, proposals):
    """ragged per-image proposals -> padded [B*Rmax] rows + counts (only copies when the lengths differ)"""
    n = [int(p.shape[0]) for p in proposals]
    B, R = len(n), max(n) if n else 0
    dev = class_logits.device
    rois = torch.zeros((B, R, 5), dtype=torch.float32, device=dev)
    for b, p in enumerate(proposals):
        rois[b, :n[b], 0] = b
        rois[b, :n[b], 1:] = p.float()
    if all(k == R for k in n):
        return class_logits, box_regression, rois.view(B * R, 5), None, B, n
    lg = torch.zeros((B, R, class_logits.shape[-1]), dtype=torch.float32, device=dev)
    rg = torch.zeros((B, R, box_regression.shape[-1]), dtype=torch.float32, device=dev)
    o = 0
    for b in range(B):
        lg[b, :n[b]] = class_logits[o:o + n[b]]
        rg[b, :n[b]] = box_regression[o:o + n[b]]
        o += n[b]
    cnt = torch.tensor(n, dtype=torch.int32, device=dev)
    return lg.view(B * R, -1), rg.view(B * R, -1), rois.view(B * R, 5), cnt, B, n


def postprocess_detections(class_logits, box_regression, proposals, image_shapes, score_thresh=0.05, nms_thresh=0.5,
                           detections_per_img=100, weights=(10.0, 10.0, 5.0, 5.0), return_ids=False, **variant):
    """torchvision RoIHeads.postprocess_detections signature -> (boxes, scores, labels) lists (one host sync for the counts)."""
    _lib.require_cuda(class_logits, box_regression, *proposals)
    shapes = [(float(h), float(w)) for (h, w) in image_shapes]
    out_b, out_s, out_l, out_i = [], [], [], []
    groups = [list(range(len(proposals)))] if len(set(shapes)) <= 1 else [[b] for b in range(len(proposals))]
    n = [int(p.shape[0]) for p in proposals]
    starts = [sum(n[:b]) for b in range(len(n))]
    for g in groups:
        if len(g) == len(proposals):
            lg, rg, props = class_logits, box_regression, list(proposals)
        else:
            b = g[0]
            lg, rg, props = class_logits[starts[b]:starts[b] + n[b]], box_regression[starts[b]:starts[b] + n[b]], [proposals[b]]
        if sum(int(p.shape[0]) for p in props) == 0:
            for _ in g:
                e = torch.zeros((0, 4), device=class_logits.device)
                out_b.append(e); out_s.append(e[:, 0]); out_l.append(e[:, 0].long()); out_i.append(e[:, 0].long())
            continue
        lgp, rgp, rois, cnt, B, _ = _pad_lists(lg, rg, props)
        pp = RoIHeadPostprocessor(shapes[g[0]], score_thresh, nms_thresh, detections_per_img, weights, **variant)
        det, idx, count = pp(lgp, rgp, rois, cnt, B=B)
        cl = count.tolist()
        for j in range(B):
            k = cl[j]
            out_b.append(det[j, :k, :4].clone()); out_s.append(det[j, :k, 4].clone()); out_l.append(det[j, :k, 5].long()); out_i.append(idx[j, :k].clone())
    return (out_b, out_s, out_l, out_i) if return_ids else (out_b, out_s, out_l)


class DecodeBox:
    """lineage call: forward(roi_cls_locs [B,R,4C], roi_scores [B,R,C], rois [B,R,4], input_shape (h,w), nms_iou=0.3, confidence=0.5)
    -> per image Tensor [k,6] = (x1,y1,x2,y2,score,label c-1) in input-image pixels, grouped by class (class order), score desc inside."""

    def __init__(self, num_classes, std=(0.1, 0.1, 0.2, 0.2)):
        self.num_classes, self.std = num_classes + 1, std

    def forward(self, roi_cls_locs, roi_scores, rois, input_shape, nms_iou=0.3, confidence=0.5):
        B, R = roi_scores.shape[0], roi_scores.shape[1]
        r5 = torch.cat((torch.arange(B, device=rois.device, dtype=torch.float32).view(B, 1, 1).expand(B, R, 1), rois.float()), 2).reshape(B * R, 5)
        pp = RoIHeadPostprocessor(input_shape, confidence, nms_iou, R * (self.num_classes - 1), self.std, mul_std=True, clamp=None,
                                  min_size=None, label_minus1=True)
        det, idx, count = pp(roi_scores.reshape(B * R, -1), roi_cls_locs.reshape(B * R, -1), r5, None, B=B)
        out = []
        for b, k in enumerate(count.tolist()):
            d = det[b, :k]
            order = torch.sort(d[:, 5], stable=True)[1]       # class-major; the score order inside a class is kept
            out.append(d[order].clone())
        return out


_META_CACHE = {}


def letterbox_meta(img1_shape, img0_shapes, device):
    """[B,5] (pad_x, pad_y, gain, w0, h0) of ultralytics scale_coords for a letterboxed input of img1_shape=(h,w); cached per
    (shapes, device) so that repeated calls do not upload again"""
    key = (tuple(img1_shape), tuple(tuple(x) for x in img0_shapes), str(device))
    hit = _META_CACHE.get(key)
    if hit is not None:
        return hit
    if len(_META_CACHE) > 64:
        _META_CACHE.clear()
    t = _letterbox_meta(img1_shape, img0_shapes, device)
    _META_CACHE[key] = t
    return t


def _letterbox_meta(img1_shape, img0_shapes, device):
    rows = []
    for (h0, w0) in img0_shapes:
        gain = min(img1_shape[0] / h0, img1_shape[1] / w0)
        rows.append(((img1_shape[1] - w0 * gain) / 2, (img1_shape[0] - h0 * gain) / 2, gain, float(w0), float(h0)))
    return torch.tensor(rows, dtype=torch.float32, device=device)


@_lib.on_device
def scale_coords(img1_shape, det, img0_shapes, count=None, xywh=False, out=None):
    """padded detections [B,max_det,6] in letterboxed pixels -> original-image pixels (clipped); xywh=True gives COCO boxes"""
    _lib.require_cuda(det, count)
    det = _lib.f32c(det)
    B, M = det.shape[0], det.shape[1]
    meta = letterbox_meta(img1_shape, img0_shapes, det.device)
    out = torch.empty_like(det) if out is None else out
    if count is not None:
        count = count.to(torch.int32).contiguous()
    _lib.check(_lib.lib().hd_scale_detections(_lib.ptr(det), _lib.ptr(count), B, M, _lib.ptr(meta), _lib.BOX_XYWH if xywh else 0, _lib.ptr(out),
                                              _lib.stream()))
    return out


def coco_records(det_xywh, count, image_ids, category_ids=None):
    """COCO result-json rows from padded (x,y,w,h,score,label) detections: one device->host copy, then host formatting"""
    d, c = det_xywh.cpu(), count.cpu().tolist()
    recs = []
    for b, k in enumerate(c):
        for r in range(k):
            x, y, w, h, s, l = d[b, r].tolist()
            cid = int(l) if category_ids is None else category_ids[int(l)]
            recs.append({"image_id": image_ids[b], "category_id": cid, "bbox": [round(x, 3), round(y, 3), round(w, 3), round(h, 3)], "score": round(s, 5)})
    return recs

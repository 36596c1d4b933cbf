"""YOLOv5 head decode + confidence filter + class-aware NMS oracle.

Reference feature: README.md:9 (YOLOv5, PAFPN head).  The dev-branch source is
not mounted; semantics follow the lineage the README credits (README.md:158-162,
ultralytics/yolov5 ``Detect.forward`` / ``non_max_suppression`` and
bubbliiiing/yolov5-v6.1-pytorch ``DecodeBox``), written out in SURVEY.md App. A.
"""
import torch
from . import boxes as _boxes

DEFAULT_ANCHORS = (
    ((10, 13), (16, 30), (33, 23)),
    ((30, 61), (62, 45), (59, 119)),
    ((116, 90), (156, 198), (373, 326)),
)
DEFAULT_STRIDES = (8, 16, 32)


def decode_box(outputs, anchors=DEFAULT_ANCHORS, strides=DEFAULT_STRIDES):
    """A.1: list of [B, A*(5+nc), H, W] -> [B, sum(A*H*W), 5+nc] = (cx,cy,w,h,obj,cls..) in px.

    view [B,A,5+nc,H,W] -> permute [B,A,H,W,5+nc]; p = sigmoid(raw);
    cx=(2p-0.5+j)*s, cy=(2p-0.5+i)*s, w=(2p)^2*aw, h=(2p)^2*ah; flatten (a,i,j)."""
    outs = []
    for x, anc, s in zip(outputs, anchors, strides):
        x = x.detach().cpu().float()
        B, C, H, W = x.shape
        A = len(anc)
        no = C // A
        p = x.view(B, A, no, H, W).permute(0, 1, 3, 4, 2).contiguous().sigmoid()
        gy, gx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
        grid = torch.stack((gx, gy), -1).view(1, 1, H, W, 2)
        anc_t = torch.tensor(anc, dtype=torch.float32).view(1, A, 1, 1, 2)
        xy = (p[..., 0:2] * 2 - 0.5 + grid) * float(s)
        wh = (p[..., 2:4] * 2) ** 2 * anc_t
        outs.append(torch.cat((xy, wh, p[..., 4:]), -1).view(B, A * H * W, no))
    return torch.cat(outs, 1)


def xywh2xyxy(x):
    y = x.clone()
    y[..., 0] = x[..., 0] - x[..., 2] / 2
    y[..., 1] = x[..., 1] - x[..., 3] / 2
    y[..., 2] = x[..., 0] + x[..., 2] / 2
    y[..., 3] = x[..., 1] + x[..., 3] / 2
    return y


def filter_candidates(pred_img, conf_thres, ge=False):
    """A.2 filter for one image [N, 5+nc] -> (cand [n,6] = xyxy,conf,cls ; anchor index [n]), anchor order.

    conf = obj * cls (fp32 product), best class = first index of the max product
    (torch.max), keep conf > thr (``ge``: >=, bubbliiiing variant)."""
    x = pred_img
    obj = x[:, 4]
    xc = (obj >= conf_thres) if ge else (obj > conf_thres)
    idx = torch.nonzero(xc).flatten()
    x = x[xc]
    if x.shape[0] == 0:
        return x.new_zeros((0, 6)), idx
    cls = x[:, 5:] * x[:, 4:5]
    box = xywh2xyxy(x[:, :4])
    conf, j = cls.max(1, keepdim=True)
    m = (conf.view(-1) >= conf_thres) if ge else (conf.view(-1) > conf_thres)
    return torch.cat((box, conf, j.float()), 1)[m], idx[m]


def filter_candidates_multi_label(pred_img, conf_thres, ge=False):
    """ultralytics multi_label=True: ``x[:, 5:] *= x[:, 4:5]``; ``i, j = (x[:, 5:] > conf_thres).nonzero().T``;
    candidates = (box[i], x[i, 5 + j], j) in (anchor, class) order.  Returned index = anchor * nc + class."""
    x = pred_img
    obj = x[:, 4]
    xc = (obj >= conf_thres) if ge else (obj > conf_thres)
    idx = torch.nonzero(xc).flatten()
    x = x[xc]
    nc = x.shape[1] - 5
    if x.shape[0] == 0:
        return x.new_zeros((0, 6)), idx
    cls = x[:, 5:] * x[:, 4:5]
    box = xywh2xyxy(x[:, :4])
    hit = (cls >= conf_thres) if ge else (cls > conf_thres)
    i, j = hit.nonzero(as_tuple=False).T
    return torch.cat((box[i], cls[i, j, None], j[:, None].float()), 1), idx[i] * nc + j


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, agnostic=False, max_det=300,
                        max_nms=30000, max_wh=7680.0, class_mode="offset", ge=False, return_index=False, multi_label=False):
    """A.2: per image filter -> (cap max_nms by conf) -> class-aware NMS -> first max_det.

    class_mode "offset": nms(boxes + cls*max_wh) (ultralytics);
               "exact" : per-class NMS on unshifted boxes (torchvision _batched_nms_vanilla).
    Returns list of [k,6]; with return_index also the anchor indices of the kept rows."""
    prediction = prediction.detach().cpu().float()
    out, out_idx = [], []
    for xi in range(prediction.shape[0]):
        x, aidx = (filter_candidates_multi_label if multi_label else filter_candidates)(prediction[xi], conf_thres, ge)
        n = x.shape[0]
        if n > max_nms:
            o = torch.sort(x[:, 4], descending=True, stable=True)[1][:max_nms]
            x, aidx = x[o], aidx[o]
        if n == 0:
            out.append(x.new_zeros((0, 6)))
            out_idx.append(aidx)
            continue
        if agnostic:
            i = _boxes.nms(x[:, :4], x[:, 4], iou_thres)
        elif class_mode == "offset":
            i = _boxes.batched_nms_offset(x[:, :4], x[:, 4], x[:, 5], iou_thres, max_wh)
        else:
            i = _boxes.batched_nms(x[:, :4], x[:, 4], x[:, 5].long(), iou_thres)
        i = i[:max_det]
        out.append(x[i])
        out_idx.append(aidx[i])
    return (out, out_idx) if return_index else out


def non_max_suppression_per_class(prediction, conf_thres=0.5, nms_thres=0.4):
    """bubbliiiing-lineage form of A.2 (yolov5-v6.1-pytorch DecodeBox.non_max_suppression): keep conf >= conf_thres, then a Python
    loop over the unique class ids (ascending), nms per class, results concatenated class by class (no max_det cut)."""
    import torchvision
    prediction = prediction.detach().cpu().float()
    out = []
    for xi in range(prediction.shape[0]):
        x, _ = filter_candidates(prediction[xi], conf_thres, ge=True)
        if x.shape[0] == 0:
            out.append(x.new_zeros((0, 6)))
            continue
        rows = []
        for c in x[:, 5].unique():
            d = x[x[:, 5] == c]
            keep = torchvision.ops.nms(d[:, :4], d[:, 4], nms_thres)
            rows.append(d[keep])
        out.append(torch.cat(rows, 0))
    return out

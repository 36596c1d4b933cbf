"""RPN proposal creation oracle.

Reference feature: README.md:8,63-65 (FasterRCNN-FPN, RPN -> RoI head).  Source
not mounted; lineage = bubbliiiing/faster-rcnn-pytorch ``ProposalCreator`` /
``loc2bbox`` / ``generate_anchor_base`` / ``_enumerate_shifted_anchor``
(README.md:158), cross-checked with torchvision
models/detection/rpn.py:231-297 and _utils.py:183-224.
"""
import math
import numpy as np
import torch
from . import boxes as _boxes


def generate_anchor_base(base_size, ratios=(0.5, 1.0, 2.0), scales=(8.0,)):
    """[len(ratios)*len(scales), 4] fp32 (x1,y1,x2,y2) centred on 0; index r*len(scales)+s.
    h = base*scale*sqrt(ratio), w = base*scale*sqrt(1/ratio)."""
    out = np.zeros((len(ratios) * len(scales), 4), np.float32)
    for i, r in enumerate(ratios):
        for j, sc in enumerate(scales):
            h = base_size * sc * math.sqrt(r)
            w = base_size * sc * math.sqrt(1.0 / r)
            out[i * len(scales) + j] = (-w / 2.0, -h / 2.0, w / 2.0, h / 2.0)
    return out


def enumerate_shifted_anchor(anchor_base, stride, H, W):
    """anchor[(i*W+j)*A+a] = anchor_base[a] + (j*stride, i*stride, j*stride, i*stride), fp32 add."""
    sx = np.arange(0, W * stride, stride, dtype=np.float32)
    sy = np.arange(0, H * stride, stride, dtype=np.float32)
    sx, sy = np.meshgrid(sx, sy)
    shift = np.stack((sx.ravel(), sy.ravel(), sx.ravel(), sy.ravel()), 1)
    A = anchor_base.shape[0]
    return (anchor_base.reshape(1, A, 4) + shift.reshape(-1, 1, 4)).reshape(-1, 4).astype(np.float32)


def loc2bbox(src_bbox, loc, clamp_dwh=None, exact_math=False):
    """bubbliiiing loc2bbox.  clamp_dwh: torchvision's bbox_xform_clip log(1000/16) (_utils.py:207-208).
    exact_math: exp evaluated in fp64 and rounded once to fp32 (the correctly rounded fp32 value): removes the libm/SIMD
    dependence of the last bit, which is what otherwise makes an end-to-end index comparison statistical."""
    w_a = src_bbox[:, 2] - src_bbox[:, 0]
    h_a = src_bbox[:, 3] - src_bbox[:, 1]
    cx_a = src_bbox[:, 0] + 0.5 * w_a
    cy_a = src_bbox[:, 1] + 0.5 * h_a
    dx, dy, dw, dh = loc[:, 0], loc[:, 1], loc[:, 2], loc[:, 3]
    if clamp_dwh is not None:
        dw = torch.clamp(dw, max=clamp_dwh)
        dh = torch.clamp(dh, max=clamp_dwh)
    cx = dx * w_a + cx_a
    cy = dy * h_a + cy_a
    if exact_math:
        w = torch.exp(dw.double()).float() * w_a
        h = torch.exp(dh.double()).float() * h_a
    else:
        w = torch.exp(dw) * w_a
        h = torch.exp(dh) * h_a
    return torch.stack((cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h), 1)


def flatten_head(objectness, deltas, score_mode="sigmoid", exact_math=False):
    """Per level NCHW head -> per-image flat (loc [B,HWA,4], fg score [B,HWA]).
    objectness [B,A,H,W] (sigmoid) or [B,2A,H,W] (softmax, channel a*2+{bg,fg});
    deltas [B,4A,H,W] channel a*4+k; flat index (i*W+j)*A+a (permute(0,2,3,1))."""
    B, C4, H, W = deltas.shape
    A = C4 // 4
    loc = deltas.permute(0, 2, 3, 1).contiguous().view(B, -1, 4)
    if score_mode == "softmax":
        s = objectness.permute(0, 2, 3, 1).contiguous().view(B, -1, 2)
        if exact_math:   # exp(x - max) / sum in fp64, rounded once
            s = s.double()
            e = torch.exp(s - s.max(-1, keepdim=True)[0])
            fg = (e[:, :, 1] / (e[:, :, 0] + e[:, :, 1])).float()
        else:
            fg = torch.softmax(s, dim=-1)[:, :, 1]
    else:
        x = objectness.permute(0, 2, 3, 1).contiguous().view(B, -1)
        fg = (1.0 / (1.0 + torch.exp(-x.double()))).float() if exact_math else x.sigmoid()
    return loc, fg


class ProposalCreator:
    """decode -> clip -> min-size (>=) -> stable sort desc -> top n_pre -> nms -> top n_post.

    The lineage pads short results by random re-sampling (np.random.choice) -- that is
    nondeterministic, so the oracle (and the product) return the unpadded rois."""

    def __init__(self, nms_iou=0.7, n_pre_nms=12000, n_post_nms=2000, min_size=16, clamp_dwh=None, exact_math=False):
        self.nms_iou, self.n_pre_nms, self.n_post_nms = nms_iou, n_pre_nms, n_post_nms
        self.min_size, self.clamp_dwh, self.exact_math = min_size, clamp_dwh, exact_math

    def __call__(self, loc, score, anchor, img_size, scale=1.0, return_index=False):
        loc, score, anchor = loc.cpu().float(), score.cpu().float(), torch.as_tensor(anchor).float()
        roi = loc2bbox(anchor, loc, self.clamp_dwh, self.exact_math)
        roi[:, [0, 2]] = torch.clamp(roi[:, [0, 2]], min=0, max=float(img_size[1]))
        roi[:, [1, 3]] = torch.clamp(roi[:, [1, 3]], min=0, max=float(img_size[0]))
        min_size = self.min_size * scale
        keep = torch.where(((roi[:, 2] - roi[:, 0]) >= min_size) & ((roi[:, 3] - roi[:, 1]) >= min_size))[0]
        roi, sc = roi[keep], score[keep]
        order = torch.sort(sc, descending=True, stable=True)[1]
        if self.n_pre_nms > 0:
            order = order[: self.n_pre_nms]
        roi, sc, src = roi[order], sc[order], keep[order]
        k = _boxes.nms(roi, sc, self.nms_iou)[: self.n_post_nms]
        if return_index:
            return roi[k], sc[k], src[k]
        return roi[k]


def rpn_proposals(objectness, deltas, anchor_bases, strides, img_size, score_mode="sigmoid", exact_math=False, **kw):
    """Multi-level, batched: lists of per-level NCHW heads -> list (per image) of rois [k,4]."""
    locs, fgs, ancs = [], [], []
    for o, d, ab, s in zip(objectness, deltas, anchor_bases, strides):
        l, f = flatten_head(o.cpu().float(), d.cpu().float(), score_mode, exact_math)
        locs.append(l)
        fgs.append(f)
        ancs.append(torch.from_numpy(enumerate_shifted_anchor(np.asarray(ab, np.float32), s, d.shape[2], d.shape[3])))
    loc, fg, anc = torch.cat(locs, 1), torch.cat(fgs, 1), torch.cat(ancs, 0)
    pc = ProposalCreator(exact_math=exact_math, **kw)
    return [pc(loc[b], fg[b], anc, img_size, return_index=True) for b in range(loc.shape[0])]

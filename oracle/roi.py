"""RoIAlign / RoIPool / FPN level assignment oracle.

Reference feature: README.md:65 (RoIAlign default), :73 (RoIPooling variant),
:73-78 ("P2" single-level vs multi-level heads).  Primary = the torchvision CPU
ops (roi_align.py:204-260, roi_pool.py:15-53, poolers.py:73-84,147-227).
"""
import math
import numpy as np
import torch
import torchvision


def roi_align(input, rois, output_size, spatial_scale=1.0, sampling_ratio=-1, aligned=False):
    return torchvision.ops.roi_align(input.cpu().float(), rois.cpu().float(), output_size, spatial_scale, sampling_ratio, aligned)


def roi_pool(input, rois, output_size, spatial_scale=1.0):
    return torchvision.ops.roi_pool(input.cpu().float(), rois.cpu().float(), output_size, spatial_scale)


def _bilinear(feat, y, x):
    """C++/CUDA kernel rule (not the Python port): sample outside [-1,H]x[-1,W] contributes 0."""
    C, H, W = feat.shape
    if y < -1.0 or y > H or x < -1.0 or x > W:
        return np.zeros(C, np.float32)
    y = max(y, np.float32(0))
    x = max(x, np.float32(0))
    y_low, x_low = int(y), int(x)
    if y_low >= H - 1:
        y_high = y_low = H - 1
        y = np.float32(y_low)
    else:
        y_high = y_low + 1
    if x_low >= W - 1:
        x_high = x_low = W - 1
        x = np.float32(x_low)
    else:
        x_high = x_low + 1
    ly = np.float32(y - np.float32(y_low))
    lx = np.float32(x - np.float32(x_low))
    hy = np.float32(1) - ly
    hx = np.float32(1) - lx
    return (hy * hx) * feat[:, y_low, x_low] + (hy * lx) * feat[:, y_low, x_high] + \
        (ly * hx) * feat[:, y_high, x_low] + (ly * lx) * feat[:, y_high, x_high]


def roi_align_restated(input, rois, output_size, spatial_scale=1.0, sampling_ratio=-1, aligned=False):
    """Scalar fp32 restatement of the C++ roi_align forward (small cases only)."""
    x = np.asarray(input, np.float32)
    rois = np.asarray(rois, np.float32)
    PH, PW = (output_size, output_size) if isinstance(output_size, int) else output_size
    K, C = rois.shape[0], x.shape[1]
    out = np.zeros((K, C, PH, PW), np.float32)
    f32 = np.float32
    off = f32(0.5) if aligned else f32(0)
    for k in range(K):
        b = int(rois[k, 0])
        sw = rois[k, 1] * f32(spatial_scale) - off
        sh = rois[k, 2] * f32(spatial_scale) - off
        ew = rois[k, 3] * f32(spatial_scale) - off
        eh = rois[k, 4] * f32(spatial_scale) - off
        rw, rh = ew - sw, eh - sh
        if not aligned:
            rw, rh = max(rw, f32(1)), max(rh, f32(1))
        bh, bw = rh / f32(PH), rw / f32(PW)
        gh = sampling_ratio if sampling_ratio > 0 else int(math.ceil(rh / f32(PH)))
        gw = sampling_ratio if sampling_ratio > 0 else int(math.ceil(rw / f32(PW)))
        count = f32(max(gh * gw, 1))
        for ph in range(PH):
            for pw in range(PW):
                acc = np.zeros(C, np.float32)
                for iy in range(gh):
                    y = sh + f32(ph) * bh + (f32(iy) + f32(0.5)) * bh / f32(gh)
                    for ix in range(gw):
                        xx = sw + f32(pw) * bw + (f32(ix) + f32(0.5)) * bw / f32(gw)
                        acc += _bilinear(x[b], f32(y), f32(xx))
                out[k, :, ph, pw] = acc / count
    return out


def roi_pool_restated(input, rois, output_size, spatial_scale=1.0):
    """round(x*scale); roi_w=max(x2-x1+1,1); floor/ceil bins clipped; empty bin -> 0."""
    x = np.asarray(input, np.float32)
    rois = np.asarray(rois, np.float32)
    PH, PW = (output_size, output_size) if isinstance(output_size, int) else output_size
    K, C, H, W = rois.shape[0], x.shape[1], x.shape[2], x.shape[3]
    out = np.zeros((K, C, PH, PW), np.float32)
    f32 = np.float32

    def rnd(v):  # C round(): half away from zero
        return int(math.floor(abs(float(v)) + 0.5) * (1 if v >= 0 else -1))
    for k in range(K):
        b = int(rois[k, 0])
        x1, y1, x2, y2 = (rnd(rois[k, i] * f32(spatial_scale)) for i in range(1, 5))
        rw, rh = max(x2 - x1 + 1, 1), max(y2 - y1 + 1, 1)
        bh, bw = f32(rh) / f32(PH), f32(rw) / f32(PW)
        for ph in range(PH):
            hs = min(max(int(math.floor(f32(ph) * bh)) + y1, 0), H)
            he = min(max(int(math.ceil(f32(ph + 1) * bh)) + y1, 0), H)
            for pw in range(PW):
                ws = min(max(int(math.floor(f32(pw) * bw)) + x1, 0), W)
                we = min(max(int(math.ceil(f32(pw + 1) * bw)) + x1, 0), W)
                if he <= hs or we <= ws:
                    continue
                out[k, :, ph, pw] = x[b, :, hs:he, ws:we].reshape(C, -1).max(1)
    return out


def level_map(rois_xyxy, k_min=2, k_max=5, canonical_scale=224.0, canonical_level=4, eps=1e-6, style="torchvision"):
    """torchvision LevelMapper (poolers.py:73-84): floor(lvl0 + log2(sqrt(area)/s0) + eps) clamp -> -k_min.
    style "mmdet": floor(log2(sqrt(area)/finest_scale + eps)) clamp [0, L-1], finest_scale=canonical_scale/4=56."""
    b = rois_xyxy.cpu().float()
    s = torch.sqrt((b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]))
    if style == "mmdet":
        lv = torch.floor(torch.log2(s / (canonical_scale / 2 ** (canonical_level - k_min)) + eps))
        return lv.clamp(min=0, max=k_max - k_min).to(torch.int64)
    lv = torch.floor(canonical_level + torch.log2(s / canonical_scale) + torch.tensor(eps, dtype=s.dtype))
    return (torch.clamp(lv, min=k_min, max=k_max).to(torch.int64) - k_min).to(torch.int64)


def multilevel_roi_align(features, rois, output_size, spatial_scales, sampling_ratio=2, aligned=False,
                         op="align", levels=None, **map_kw):
    """poolers.py:147-227: assign level, pool each RoI on its level, keep original RoI order.
    features: list of [B,C,H_l,W_l]; rois [K,5]; single level (the README "P2" variant) if len==1."""
    rois = rois.cpu().float()
    K, C = rois.shape[0], features[0].shape[1]
    PH, PW = (output_size, output_size) if isinstance(output_size, int) else output_size
    if levels is None:
        levels = level_map(rois[:, 1:5], **map_kw) if len(features) > 1 else torch.zeros(K, dtype=torch.int64)
    out = torch.zeros((K, C, PH, PW), dtype=torch.float32)
    for l, (f, sc) in enumerate(zip(features, spatial_scales)):
        idx = torch.where(levels == l)[0]
        if idx.numel() == 0:
            continue
        if op == "align":
            out[idx] = roi_align(f, rois[idx], (PH, PW), sc, sampling_ratio, aligned)
        else:
            out[idx] = roi_pool(f, rois[idx], (PH, PW), sc)
    return out, levels

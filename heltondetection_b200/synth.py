"""Synthetic head tensors for the five BASELINE.json configs.

"Planted objects": background logits are Gaussian noise; for G ground-truth boxes
per image the matching (level, anchor, cell) entries are overwritten with the
*inverse* of the decode so that clusters of anchors regress to the same box
(+ jitter) -- this is what gives NMS/WBF realistic overlap structure.  Everything
is generated on the CPU with a seeded ``torch.Generator`` / ``numpy`` RNG and then
copied, so the CPU oracle and the CUDA path see identical bits.
"""
import math
import numpy as np
import torch

YOLO_ANCHORS = (
    ((10, 13), (16, 30), (33, 23)),
    ((30, 61), (62, 45), (59, 119)),
    ((116, 90), (156, 198), (373, 326)),
)
YOLO_STRIDES = (8, 16, 32)
LOGIT_CLAMP = 12.0


def _logit(p):
    return math.log(p / (1.0 - p))


def sample_gt(rng, G, img, dense=False):
    """G boxes (cx,cy,w,h) px + class, COCO-like (sides 16..0.6*img) or dense (8..48 px)."""
    lo, hi = (8.0, 48.0) if dense else (16.0, 0.6 * img)
    w = np.exp(rng.uniform(math.log(lo), math.log(hi), G))
    h = w * np.exp(rng.uniform(-0.7, 0.7, G))
    h = np.clip(h, lo / 2, img * 0.9)
    cx = rng.uniform(0.05 * img, 0.95 * img, G)
    cy = rng.uniform(0.05 * img, 0.95 * img, G)
    return np.stack((cx, cy, w, h), 1)


def plant_yolo(heads, b, gt, cls, rng, nc, anchors=YOLO_ANCHORS, strides=YOLO_STRIDES, max_anchors=2):
    """Overwrite head logits of image b so anchors near each gt decode to it (inverse of A.1)."""
    flat = [(l, a, aw, ah) for l, lv in enumerate(anchors) for a, (aw, ah) in enumerate(lv)]
    no = 5 + nc
    heads = [t.numpy() if isinstance(t, torch.Tensor) else t for t in heads]   # views: element writes without dispatcher overhead
    for (cx, cy, w, h), c in zip(gt, cls):
        ratios = [max(w / aw, aw / w, h / ah, ah / h) for (_, _, aw, ah) in flat]
        order = np.argsort(ratios)[:max_anchors]
        for k in order:
            if ratios[k] >= 3.9:
                continue
            l, a, aw, ah = flat[k]
            s = strides[l]
            t = heads[l]
            H, W = t.shape[2], t.shape[3]
            j0, i0 = int(cx / s), int(cy / s)
            for di in (-1, 0, 1):
                for dj in (-1, 0, 1):
                    i, j = i0 + di, j0 + dj
                    if not (0 <= i < H and 0 <= j < W):
                        continue
                    ox, oy = cx / s - j, cy / s - i
                    if not (-0.45 < ox < 1.45 and -0.45 < oy < 1.45):
                        continue
                    px, py = (ox + 0.5) / 2, (oy + 0.5) / 2
                    pw = min(max(math.sqrt(w / aw) / 2, 0.02), 0.98)
                    ph = min(max(math.sqrt(h / ah) / 2, 0.02), 0.98)
                    jit = rng.normal(0, 0.05, 4)
                    base = a * no
                    t[b, base + 0, i, j] = _logit(px) + jit[0]
                    t[b, base + 1, i, j] = _logit(py) + jit[1]
                    t[b, base + 2, i, j] = _logit(pw) + jit[2]
                    t[b, base + 3, i, j] = _logit(ph) + jit[3]
                    t[b, base + 4, i, j] = rng.normal(2.0, 1.0)
                    t[b, base + 5 + int(c), i, j] = rng.normal(2.0, 1.0)


def yolo_background(B, img, nc, gen, strides=YOLO_STRIDES, A=3):
    heads = []
    no = 5 + nc
    for s in strides:
        H = W = img // s
        x = torch.randn((B, A, no, H, W), generator=gen, dtype=torch.float32)
        x[:, :, 4].mul_(1.5).sub_(7.0)
        x[:, :, 5:].sub_(4.0)
        heads.append(x.view(B, A * no, H, W))
    return heads


def yolo_heads(B, img=640, nc=80, G=20, seed=1234, dense=False, gts=None):
    """-> (list of 3 heads [B, 3*(5+nc), img/s, img/s] fp32 CPU, list of per-image gt arrays)."""
    gen = torch.Generator().manual_seed(seed)
    rng = np.random.default_rng(seed)
    heads = yolo_background(B, img, nc, gen)
    out_gt = []
    for b in range(B):
        if gts is None:
            gt = sample_gt(rng, G, img, dense)
            cls = rng.integers(0, nc, len(gt))
        else:
            gt, cls = gts[b]
        plant_yolo(heads, b, gt, cls, rng, nc)
        out_gt.append((gt, cls))
    for t in heads:
        t.clamp_(-LOGIT_CLAMP, LOGIT_CLAMP)
    return heads, out_gt


# ----------------------------------------------------------------------------- FasterRCNN
RPN_STRIDES = (4, 8, 16, 32)
RPN_RATIOS = (0.5, 1.0, 2.0)
RPN_SCALE = 8.0


def rpn_anchor_bases(strides=RPN_STRIDES, ratios=RPN_RATIOS, scale=RPN_SCALE):
    """Per level [A,4] fp32 base anchors centred on 0 (h = s*scale*sqrt(r), w = s*scale/sqrt(r))."""
    out = []
    for s in strides:
        ab = np.zeros((len(ratios), 4), np.float32)
        for i, r in enumerate(ratios):
            h = s * scale * math.sqrt(r)
            w = s * scale * math.sqrt(1.0 / r)
            ab[i] = (-w / 2.0, -h / 2.0, w / 2.0, h / 2.0)
        out.append(ab)
    return out


def rpn_heads(B, img=832, G=20, seed=1237, softmax=False, strides=RPN_STRIDES):
    """-> objectness list [B,A or 2A,H,W], deltas list [B,4A,H,W], anchor bases, gts."""
    gen = torch.Generator().manual_seed(seed)
    rng = np.random.default_rng(seed)
    A = len(RPN_RATIOS)
    bases = rpn_anchor_bases(strides)
    obj, dlt = [], []
    for s in strides:
        H = W = img // s
        o = torch.randn((B, A * (2 if softmax else 1), H, W), generator=gen) * 1.5 - (0.0 if softmax else 4.0)
        if softmax:
            o.view(B, A, 2, H, W)[:, :, 0].add_(2.0)
            o.view(B, A, 2, H, W)[:, :, 1].sub_(2.0)
        d = torch.randn((B, 4 * A, H, W), generator=gen) * 0.3
        obj.append(o)
        dlt.append(d)
    gts = []
    for b in range(B):
        gt = sample_gt(rng, G, img, False)
        gts.append(gt)
        for cx, cy, w, h in gt:
            size = math.sqrt(w * h)
            l = int(np.argmin([abs(math.log(size / (s * RPN_SCALE))) for s in strides]))
            s = strides[l]
            H = W = img // s
            a = int(np.argmin([abs(math.log((h / w) / r)) for r in RPN_RATIOS]))
            ab = bases[l][a]
            aw, ah = float(ab[2] - ab[0]), float(ab[3] - ab[1])
            j0, i0 = int(round(cx / s)), int(round(cy / s))
            for di in (-1, 0, 1):
                for dj in (-1, 0, 1):
                    i, j = i0 + di, j0 + dj
                    if not (0 <= i < H and 0 <= j < W):
                        continue
                    acx, acy = j * s, i * s
                    jit = rng.normal(0, 0.03, 4)
                    dlt[l][b, a * 4 + 0, i, j] = (cx - acx) / aw + jit[0]
                    dlt[l][b, a * 4 + 1, i, j] = (cy - acy) / ah + jit[1]
                    dlt[l][b, a * 4 + 2, i, j] = math.log(w / aw) + jit[2]
                    dlt[l][b, a * 4 + 3, i, j] = math.log(h / ah) + jit[3]
                    if softmax:
                        obj[l][b, a * 2 + 1, i, j] = rng.normal(3.0, 1.0)
                        obj[l][b, a * 2 + 0, i, j] = rng.normal(-1.0, 1.0)
                    else:
                        obj[l][b, a, i, j] = rng.normal(3.0, 1.0)
    return obj, dlt, bases, gts


def fpn_features(B, img=832, C=256, seed=1237, strides=RPN_STRIDES):
    gen = torch.Generator().manual_seed(seed + 77)
    return [torch.randn((B, C, img // s, img // s), generator=gen) for s in strides]


def random_rois(B, K_per_img, img, seed=0, min_side=8.0):
    """[B*K,5] (batch_idx, x1,y1,x2,y2) px, log-uniform sizes, clipped to the image."""
    rng = np.random.default_rng(seed)
    n = B * K_per_img
    side = np.exp(rng.uniform(math.log(min_side), math.log(0.8 * img), n))
    w = side * np.exp(rng.uniform(-0.6, 0.6, n))
    h = side * np.exp(rng.uniform(-0.6, 0.6, n))
    cx, cy = rng.uniform(0, img, n), rng.uniform(0, img, n)
    x1, y1 = np.clip(cx - w / 2, 0, img), np.clip(cy - h / 2, 0, img)
    x2, y2 = np.clip(cx + w / 2, 0, img), np.clip(cy + h / 2, 0, img)
    bi = np.repeat(np.arange(B), K_per_img)
    return torch.from_numpy(np.stack((bi, x1, y1, x2, y2), 1).astype(np.float32))


# ----------------------------------------------------------------------------- TTA
TTA_VIEWS = ((544, False), (544, True), (640, False), (640, True), (768, False), (768, True))


def tta_heads(B, img=640, nc=80, G=20, seed=1239, views=TTA_VIEWS):
    """Per view: heads of the resized (+h-flipped) image containing the same gt objects.
    -> list over views of (heads, scale r, hflip, view size)."""
    rng = np.random.default_rng(seed)
    gts = []
    for _ in range(B):
        gt = sample_gt(rng, G, img, False)
        gts.append((gt, rng.integers(0, nc, len(gt))))
    out = []
    for v, (size, flip) in enumerate(views):
        r = size / img
        vg = []
        for gt, cls in gts:
            g = gt * r
            if flip:
                g = g.copy()
                g[:, 0] = size - g[:, 0]
            vg.append((g, cls))
        heads, _ = yolo_heads(B, size, nc, G, seed + 100 * (v + 1), gts=vg)
        out.append((heads, r, flip, size))
    return out, gts

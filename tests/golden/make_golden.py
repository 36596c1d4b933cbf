#!/usr/bin/env python
"""Generates tests/golden/golden_v1.npz.

The reference's own source/tests are not in the mount (README only), so the fixtures are produced by the
one executable piece of the reference's path -- the torchvision 0.26.0 CPU ops -- plus the lineage
restatements in oracle/ for the stages torchvision does not cover.  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys
import numpy as np
import torch
import torchvision

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from heltondetection_b200 import synth  # noqa: E402


def main():
    out = {"torchvision_version": np.array(torchvision.__version__), "torch_version": np.array(torch.__version__)}
    g = torch.Generator().manual_seed(42)
    # ---- nms / batched_nms / box_iou (torchvision CPU)
    n = 200
    ctr = torch.rand((25, 2), generator=g) * 300
    wh = torch.rand((25, 2), generator=g) * 60 + 10
    pick = torch.randint(0, 25, (n,), generator=g)
    c = ctr[pick] + torch.randn((n, 2), generator=g) * 3
    s = wh[pick] * (1 + 0.1 * torch.randn((n, 2), generator=g))
    boxes = torch.cat((c - s / 2, c + s / 2), 1)
    scores = torch.rand((n,), generator=g)
    cls = torch.randint(0, 4, (n,), generator=g)
    out["nms_boxes"], out["nms_scores"], out["nms_cls"] = boxes.numpy(), scores.numpy(), cls.numpy()
    for thr in (0.3, 0.5, 0.7):
        out[f"nms_keep_{thr}"] = torchvision.ops.nms(boxes, scores, thr).numpy()
    out["batched_keep_0.5"] = torchvision.ops.boxes._batched_nms_vanilla(boxes, scores, cls, 0.5).numpy()
    out["box_iou_40x50"] = torchvision.ops.box_iou(boxes[:40], boxes[40:90]).numpy()
    # ---- roi_align / roi_pool (torchvision CPU), incl. RoIs outside the map
    x = torch.randn((2, 6, 14, 12), generator=g)
    rois = synth.random_rois(2, 10, 96, 7)
    rois[0, 1:] = torch.tensor([-20.0, -10.0, 30.0, 40.0])
    rois[1, 1:] = torch.tensor([80.0, 90.0, 130.0, 150.0])
    out["roi_x"], out["roi_rois"] = x.numpy(), rois.numpy()
    for sr in (2, 0):
        for al in (False, True):
            out[f"roi_align_sr{sr}_al{int(al)}"] = torchvision.ops.roi_align(x, rois, (7, 7), 0.125, sr, al).numpy()
    out["roi_pool"] = torchvision.ops.roi_pool(x, rois, (7, 7), 0.125).numpy()
    sides = torch.tensor([5.0, 111.9, 112.0, 224.0, 448.0, 900.0])
    lb = torch.stack((torch.zeros(6), torch.zeros(6), sides, sides), 1)
    out["level_boxes"], out["level_ids"] = lb.numpy(), oracle.roi.level_map(lb).numpy()
    # ---- YOLO decode + NMS (lineage restatement over torch CPU + torchvision nms)
    heads, _ = synth.yolo_heads(1, 128, 3, 4, 11)
    pred = oracle.yolo.decode_box(heads)
    det, idx = oracle.yolo.non_max_suppression(pred, 0.25, 0.45, return_index=True)
    for l, h in enumerate(heads):
        out[f"yolo_head{l}"] = h.numpy()
    out["yolo_pred"], out["yolo_det"], out["yolo_idx"] = pred.numpy(), det[0].numpy(), idx[0].numpy()
    # ---- RPN proposals
    obj, dlt, bases, _ = synth.rpn_heads(1, 128, G=4, seed=13)
    roi, sc, ix = oracle.rpn.rpn_proposals(obj, dlt, bases, (4, 8, 16, 32), (128, 128), n_pre_nms=600, n_post_nms=100, min_size=8)[0]
    for l in range(4):
        out[f"rpn_obj{l}"], out[f"rpn_dlt{l}"], out[f"rpn_base{l}"] = obj[l].numpy(), dlt[l].numpy(), bases[l]
    out["rpn_roi"], out["rpn_score"], out["rpn_idx"] = roi.numpy(), sc.numpy(), ix.numpy()
    # ---- WBF
    rng = np.random.default_rng(3)
    bl, sl, ll = [], [], []
    gc = rng.uniform(0.2, 0.8, (5, 2)); gw = rng.uniform(0.05, 0.25, (5, 2)); gl = rng.integers(0, 3, 5)
    for v in range(3):
        m = 12
        p = rng.integers(0, 5, m)
        cc = gc[p] + rng.normal(0, 0.01, (m, 2)); ss = gw[p] * (1 + rng.normal(0, 0.05, (m, 2)))
        bl.append(np.concatenate((cc - ss / 2, cc + ss / 2), 1).astype(np.float32))
        sl.append(rng.uniform(0.05, 1, m).astype(np.float32)); ll.append(gl[p].astype(np.float32))
    fb, fs, fl = oracle.wbf.weighted_boxes_fusion(bl, sl, ll, None, 0.55, 0.1, "avg")
    for v in range(3):
        out[f"wbf_b{v}"], out[f"wbf_s{v}"], out[f"wbf_l{v}"] = bl[v], sl[v], ll[v]
    out["wbf_boxes"], out["wbf_scores"], out["wbf_labels"] = fb, fs, fl
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

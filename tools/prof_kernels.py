#!/usr/bin/env python
"""Runs every hd_b200 kernel once on BASELINE-shaped inputs (reduced batch) -- the target of the ncu captures
kept under profiles/.  usage: python tools/prof_kernels.py"""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heltondetection_b200 import synth, yolo, rpn, roi, wbf, ops  # noqa: E402

torch.cuda.set_device(0)
# cfg2 (B=256 so the decode kernel is the real bench launch)
heads, _ = synth.yolo_heads(256, 640, 80, 20, 1235)
heads = [h.cuda() for h in heads]
for dense in (True, False):
    det, cnt, _ = yolo.YoloPostprocessor(conf_thres=0.25, iou_thres=0.45, dense_read=dense)(heads)
print("cfg2 kept/img", cnt.float().mean().item())
pred = yolo.decode_box([h[:32] for h in heads])
del heads, pred
# cfg4 (dense scenes, large-image NMS path)
heads, _ = synth.yolo_heads(16, 1280, 10, 300, 1238, dense=True)
det, cnt, _ = yolo.YoloPostprocessor(conf_thres=0.001, iou_thres=0.6)([h.cuda() for h in heads])
print("cfg4 kept/img", cnt.float().mean().item())
# cfg3
B, img = 4, 832
obj, dlt, bases, _ = synth.rpn_heads(B, img, G=20, seed=1237)
feats = [f.cuda() for f in synth.fpn_features(B, img, 256, 1237)]
pr = rpn.RpnProposals(bases, (4, 8, 16, 32), (img, img), n_pre_nms=12000, n_post_nms=2000, min_size=16)
rois, cnt, sc, idx = pr([o.cuda() for o in obj], [d.cuda() for d in dlt])
scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
out, lv = roi.multilevel_roi_align(feats, rois, 7, scales, 2, False)                  # layout pass + quad kernel
out, lv = roi.multilevel_roi_align(feats, rois, 7, scales, 2, False, op="pool", layout="nhwc")
print("cfg3 proposals/img", cnt.tolist(), out.shape)
iou = ops.box_iou(rois[:2000, 1:], rois[2000:4000, 1:])
# cfg5
views, _ = synth.tta_heads(8, 640, 80, G=20, seed=1239)
fusion = wbf.TTAFusion([(r, f, s) for (_, r, f, s) in views], (640, 640), 80, max_det=300, iou_thr=0.55, skip_box_thr=0.001)
pp = yolo.YoloPostprocessor(conf_thres=0.25, iou_thres=0.45)
for v, (h, r, f, s) in enumerate(views):
    det, cnt, _ = pp([x.cuda() for x in h])
    fusion.map_back(v, det, cnt)
ob, os_, ol, oc = fusion.fuse()
torch.cuda.synchronize()
print("cfg5 fused/img", oc.float().mean().item())

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
# next rows (8f): RoI-head post-process, label assignment, RoIAlign backward, fp16 heads
from heltondetection_b200 import roi_head, assign  # noqa: E402
g = torch.Generator().manual_seed(1240)
R = rois.shape[0] // B
lg = (torch.randn(B * R, 81, generator=g) * 3).cuda()
rg = (torch.randn(B * R, 324, generator=g) * 0.5).cuda()
det, didx, dcnt = roi_head.RoIHeadPostprocessor((img, img), 0.05, 0.5, 100)(lg, rg, rois, cnt, B=B)
sc_ = roi_head.scale_coords((img, img), det, [(1080, 1920)] * B, dcnt, xywh=True)
gtb = rois[:64, 1:].contiguous()
m = assign.Matcher(0.7, 0.3, True)(gtb, rois[:, 1:].contiguous())
nhwc0 = feats[0].contiguous(memory_format=torch.channels_last)
gi = roi.roi_align_backward(out[:2000].contiguous(), rois[:2000].contiguous(), 0.25, 7, 7, B, 256, nhwc0.shape[2], nhwc0.shape[3], 2, False)
roi.set_mode(2)
o2 = roi.roi_align(nhwc0, rois[:2000].contiguous(), 7, 0.25, 2, False)
roi.set_mode(0)
h16, _ = synth.yolo_heads(64, 640, 80, 20, 1235)
d16 = yolo.YoloPostprocessor(conf_thres=0.25, iou_thres=0.45, dense_read=True)([h.half().cuda() for h in h16])
del h16
tg = assign.encode_boxes(gtb, rois[:, 1:].contiguous(), m, (10.0, 10.0, 5.0, 5.0))
po, pam = roi.roi_pool_with_argmax(nhwc0, rois[:2000].contiguous(), 7, 0.25)
gp = roi.roi_pool_backward(po, rois[:2000].contiguous(), pam, 0.25, 7, 7, B, 256, nhwc0.shape[2], nhwc0.shape[3], channels_last=True)
hcl, _ = synth.yolo_heads(64, 640, 80, 20, 1235)
dcl = yolo.YoloPostprocessor(conf_thres=0.25, iou_thres=0.45, dense_read=True)([h.cuda().contiguous(memory_format=torch.channels_last) for h in hcl])
del hcl
print("8f rows ok", dcnt.tolist(), int((m >= 0).sum()))
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

"""Small invocation of every kernel added in the second half of round 1 (all cluster sizes, both NMS modes, every head dtype /
layout).  Written for `compute-sanitizer --tool memcheck`; that tool is closed on this GPU pool, so it serves as a quick coverage run."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heltondetection_b200 import synth, yolo, rpn, roi, ops, roi_head, assign, _lib

torch.cuda.set_device(0)
g = torch.Generator().manual_seed(0)
# RPN: cluster kernel at every cluster size, two rank batches (n_post small)
obj, dlt, bases, _ = synth.rpn_heads(2, 256, G=6, seed=1237)
obj, dlt = [o.cuda() for o in obj], [d.cuda() for d in dlt]
for cl in (8, 4, 2, 1):
    _lib.lib().hd_rpn_set_cluster_size(cl)
    pr = rpn.RpnProposals(bases, (4, 8, 16, 32), (256, 256), n_pre_nms=3000, n_post_nms=150, min_size=4)
    rois, cnt, sc, idx = pr(obj, dlt)
_lib.lib().hd_rpn_set_cluster_size(0)
# detection NMS: cluster kernel (B=3), class-aware, and the single-CTA path
n = 3000
xy = torch.rand(3, n, 2, generator=g) * 400
bx = torch.cat((xy, xy + torch.rand(3, n, 2, generator=g) * 60 + 4), 2).cuda()
ss = torch.rand(3, n, generator=g).cuda()
cc = torch.randint(0, 5, (3, n), generator=g).int().cuda()
for b in range(3):
    for mode in (0, 1):
        ops.set_nms_mode(mode)
        ops.nms(bx[b], ss[b], 0.5); ops.batched_nms(bx[b], ss[b], cc[b], 0.5)
ops.set_nms_mode(0)
# RoIAlign: gather + staged-row kernel, backward, RoIPool backward
feats = synth.fpn_features(2, 256, 64, seed=5)
x = feats[0].cuda().contiguous(memory_format=torch.channels_last)
r = synth.random_rois(2, 320, 256, 9).cuda()
for mode in (1, 2):
    roi.set_mode(mode)
    o = roi.roi_align(x, r, 7, 0.25, 2, False)
roi.set_mode(0)
gi = roi.roi_align_backward(o, r, 0.25, 7, 7, 2, 64, x.shape[2], x.shape[3], 2, False)
po, am = roi.roi_pool_with_argmax(x, r, 7, 0.25)
gp = roi.roi_pool_backward(po, r, am, 0.25, 7, 7, 2, 64, x.shape[2], x.shape[3], channels_last=True)
# YOLO: fp16 and NHWC heads (thread-loaded sparse + TMA dense)
heads, _ = synth.yolo_heads(2, 160, 20, 6, 3)
for dense in (False, True):
    pp = yolo.YoloPostprocessor(conf_thres=0.25, iou_thres=0.45, dense_read=dense)
    pp([h.half().cuda() for h in heads]); pp([h.cuda().contiguous(memory_format=torch.channels_last) for h in heads]); pp([h.cuda() for h in heads])
# RoI head, output formats, label assignment, targets
R = 200
lg = (torch.randn(2 * R, 9, generator=g) * 3).cuda(); rg = (torch.randn(2 * R, 36, generator=g) * 0.5).cuda()
rr = torch.cat((torch.arange(2).repeat_interleave(R)[:, None].float().cuda(), bx[0, :2 * R]), 1)
det, did, dc = roi_head.RoIHeadPostprocessor((400, 460))(lg, rg, rr, None, B=2)
sc2 = roi_head.scale_coords((400, 460), det, [(800, 920)] * 2, dc, xywh=True)
m = assign.Matcher(0.7, 0.3, True)(bx[0, :40].contiguous(), bx[1, :1000].contiguous())
t = assign.encode_boxes(bx[0, :40].contiguous(), bx[1, :1000].contiguous(), m, (10.0, 10.0, 5.0, 5.0))
torch.cuda.synchronize()
print("sanitize_small ok", cnt.tolist(), dc.tolist(), int((m >= 0).sum()))

"""cfg3 RoIAlign only, a few calls (target of ncu captures).  usage: python tools/roi_min.py [mode]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heltondetection_b200 import synth, rpn, roi
B, img = 16, 832
obj, dlt, bases, _ = synth.rpn_heads(B, img, G=20, seed=1237)
feats = [f.cuda().contiguous(memory_format=torch.channels_last) for f in synth.fpn_features(B, img, 256, 1237)]
pr = rpn.RpnProposals(bases, (4, 8, 16, 32), (img, img), n_pre_nms=12000, n_post_nms=2000, min_size=16)
rois, cnt, _, _ = pr([o.cuda() for o in obj], [d.cuda() for d in dlt])
roi.set_mode(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
for _ in range(4):
    out = roi.multilevel_roi_align(feats, rois, 7, [1 / 4, 1 / 8, 1 / 16, 1 / 32], 2, False)[0]
torch.cuda.synchronize()
print("ok", out.shape, float(out.abs().mean()))

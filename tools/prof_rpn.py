import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heltondetection_b200 import synth, rpn, _lib
B, img = 4, 832
obj, dlt, bases, _ = synth.rpn_heads(B, img, G=20, seed=1237)
obj, dlt = [o.cuda() for o in obj], [d.cuda() for d in dlt]
_lib.lib().hd_rpn_set_cluster_size(8)
pr = rpn.RpnProposals(bases, (4, 8, 16, 32), (img, img), n_pre_nms=12000, n_post_nms=2000, min_size=16)
for _ in range(3):
    rois, cnt, sc, idx = pr(obj, dlt)
torch.cuda.synchronize()
print("ok", cnt.tolist())

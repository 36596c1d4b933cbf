import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heltondetection_b200 import synth, rpn, roi
B, img = 16, 832
obj, dlt, bases, _ = synth.rpn_heads(B, img, G=20, seed=1237)
feats = [f.cuda().contiguous(memory_format=torch.channels_last) for f in synth.fpn_features(B, img, 256, 1237)]
pr = rpn.RpnProposals(bases, (4, 8, 16, 32), (img, img), n_pre_nms=12000, n_post_nms=2000, min_size=16)
rois, cnt, _, _ = pr([o.cuda() for o in obj], [d.cuda() for d in dlt])
scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
flush = torch.empty((256 << 20,), dtype=torch.uint8, device="cuda")
def timed(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        flush.zero_(); a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / iters
for dbg, name in ((0, "full"), (1, "no prefetch"), (2, "no row walk (prefetch only)"), (3, "no prefetch, no walk"), (7, "nothing but table + sync"), (4, "no tile store")):
    roi.set_mode(3 | (dbg << 4))
    print(f"{name:32s} {timed(lambda: roi.multilevel_roi_align(feats, rois, 7, scales, 2, False)) * 1e3:8.1f} us")
roi.set_mode(0)

#!/usr/bin/env python
"""cfg3 RoIAlign column-owner kernel, ablations through hd_roi_set_mode bits 4..7 (1: no output store, 2: no loads): where the time goes."""
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


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        flush.zero_(); a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / iters


for name, dbg in (("full", 0), ("full, thread stores (no TMA)", 4), ("no output store", 1), ("no loads", 2), ("no loads, thread stores", 6), ("no loads, no output store", 3)):
    roi.set_mode(1 | (dbg << 4))
    t = timed(lambda: roi.multilevel_roi_align(feats, rois, 7, scales, 2, False))
    print(f"{name:28s} {t * 1e3:8.1f} us", flush=True)
roi.set_mode(0)
out = torch.empty((rois.shape[0], 256, 7, 7), device="cuda")
print(f"{'torch fill of the output':28s} {timed(lambda: out.fill_(1.0)) * 1e3:8.1f} us")
print(f"{'torch empty (allocator)':28s} {timed(lambda: torch.empty((rois.shape[0], 256, 7, 7), device='cuda')) * 1e3:8.1f} us")

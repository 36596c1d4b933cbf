#!/usr/bin/env python
"""cfg3 RoIAlign gather kernel: table rows in flight per thread (1 = round-1 form, 2 = default) -- developer experiment, L2 flushed."""
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
nbytes = sum(f.numel() * 4 for f in feats) + rois.shape[0] * 256 * 49 * 4


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        flush.zero_(); a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / iters


roi.set_mode(1)
ref = roi.multilevel_roi_align(feats, rois, 7, scales, 2, False)[0]
for name, mode in (("1 row in flight (4 CTAs/SM)", 1 | (1 << 8)), ("2 rows in flight (3 CTAs/SM, default)", 1)):
    roi.set_mode(mode)
    for sr in (2, 1):
        t = timed(lambda: roi.multilevel_roi_align(feats, rois, 7, scales, sr, False))
        eq = bool(torch.equal(roi.multilevel_roi_align(feats, rois, 7, scales, 2, False)[0], ref))
        print(f"{name:34s} sr={sr}: {t * 1e3:8.1f} us  {nbytes / t / 1e6:7.1f} GB/s ({nbytes / t / 1e6 / 6536.7 * 100:.1f}% of measured peak)  bit-equal {eq}", flush=True)
roi.set_mode(0)

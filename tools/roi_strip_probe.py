#!/usr/bin/env python
"""cfg3 RoIAlign: streamed (strip) kernel vs per-RoI gather kernel, L2 flushed between iterations (developer tool)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heltondetection_b200 import synth, rpn, roi, ops

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


lv = ops.level_map(rois[:, 1:])
print("rois per level:", [int((lv == l).sum()) for l in range(4)])
w = (rois[:, 3] - rois[:, 1]); h = (rois[:, 4] - rois[:, 2])
print("roi side px: mean %.1f  p50 %.1f  p90 %.1f  max %.1f" % (float(w.mean()), float(w.median()), float(w.kthvalue(int(0.9 * w.numel()))[0]), float(w.max())))
for mode, name in ((3, "row-walk"), (2, "streamed"), (1, "gather")):
    roi.set_mode(mode)
    for sr in (2,):
        t = timed(lambda: roi.multilevel_roi_align(feats, rois, 7, scales, sr, False))
        print(f"{name:9s} sr={sr}: {t * 1e3:8.1f} us  {nbytes / t / 1e6:7.1f} GB/s ({nbytes / t / 1e6 / 6536.7 * 100:.1f}% of measured peak)")
roi.set_mode(1)
b = roi.multilevel_roi_align(feats, rois, 7, scales, 2, False)[0]
for mode in (3, 2):
    roi.set_mode(mode)
    a = roi.multilevel_roi_align(feats, rois, 7, scales, 2, False)[0]
    print("mode", mode, "bit-equal to gather:", bool(torch.equal(a, b)))
roi.set_mode(2)
import ctypes as C
from heltondetection_b200 import _lib
L = _lib.lib()
L.hd_debug_roi_profile(1, None)
roi.multilevel_roi_align(feats, rois, 7, scales, 2, False)
a8 = (C.c_ulonglong * 8)()
L.hd_debug_roi_profile(1, a8)
v = list(a8)
items = max(v[4], 1)
print("streamed kernel, per item-slice (cycles): tables %.0f  waits %.0f  rows %.0f  output %.0f | items %d | producer wait total %.1f Mcycles" %
      (v[0] / items, v[1] / items, v[2] / items, v[3] / items, v[4], v[5] / 1e6))
L.hd_debug_roi_profile(0, None)

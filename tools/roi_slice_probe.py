#!/usr/bin/env python
"""cfg3 RoIAlign: channel-sliced kernel walking a position-sorted RoI list (L1 reuse between neighbours) vs the per-RoI gather kernel.
Developer experiment: the list is sorted with torch here; L2 flushed between iterations."""
import os, sys, ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heltondetection_b200 import synth, rpn, roi, ops, _lib

B, img = 16, 832
obj, dlt, bases, _ = synth.rpn_heads(B, img, G=20, seed=1237)
feats = [f.cuda().contiguous(memory_format=torch.channels_last) for f in synth.fpn_features(B, img, 256, 1237)]
pr = rpn.RpnProposals(bases, (4, 8, 16, 32), (img, img), n_pre_nms=12000, n_post_nms=2000, min_size=16)
rois, cnt, _, _ = pr([o.cuda() for o in obj], [d.cuda() for d in dlt])
scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
flush = torch.empty((256 << 20,), dtype=torch.uint8, device="cuda")
K = rois.shape[0]
nbytes = sum(f.numel() * 4 for f in feats) + K * 256 * 49 * 4


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        flush.zero_(); a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / iters


def rep(name, t):
    print(f"{name:46s} {t * 1e3:8.1f} us  {nbytes / t / 1e6:7.1f} GB/s ({nbytes / t / 1e6 / 6536.7 * 100:.1f}% of measured peak)", flush=True)


roi.set_mode(1)
ref, lv64 = roi.multilevel_roi_align(feats, rois, 7, scales, 2, False)
rep("gather (one RoI per CTA, K CTAs)", timed(lambda: roi.multilevel_roi_align(feats, rois, 7, scales, 2, False)))
lv = lv64.to(torch.int32)
arr = roi._levels_struct(feats, scales)
for l, f in enumerate(feats):
    arr[l].H, arr[l].W = f.shape[2], f.shape[3]
L = _lib.lib()
sc = torch.tensor(scales, device="cuda")[lv64]
x0 = (rois[:, 1] * sc).floor().clamp(min=0).long(); y0 = (rois[:, 2] * sc).floor().clamp(min=0).long()
bi = rois[:, 0].long()
orders = {
    "unsorted": torch.arange(K, device="cuda"),
    "level,image,y/4,x": torch.argsort(((lv64 * 64 + bi) * 64 + y0 // 4) * 1024 + x0),
    "level,image,y/8,x": torch.argsort(((lv64 * 64 + bi) * 64 + y0 // 8) * 1024 + x0),
    "level,image,y/8,x/8,y,x": torch.argsort(((((lv64 * 64 + bi) * 64 + y0 // 8) * 64 + x0 // 8) * 1024 + y0) * 1024 + x0),
}
cnt_t = torch.tensor([K], dtype=torch.int32, device="cuda")
out = torch.empty_like(ref)
for oname, order in orders.items():
    lst = order.to(torch.int32).contiguous()
    for slices in (1, 2, 4):
        for chunk in (4, 16, 54):
            def run():
                _lib.check(L.hd_debug_roi_align_sliced(arr, 4, 256, _lib.ptr(rois), _lib.ptr(lv), K, 7, 7, 2, 0, _lib.ptr(out), _lib.ptr(lst), _lib.ptr(cnt_t),
                                                       slices, chunk, _lib.stream()))
            out.zero_(); run()
            eq = bool(torch.equal(out, ref))
            rep(f"sliced [{oname}] slices={slices} chunk={chunk} eq={eq}", timed(run, 5))

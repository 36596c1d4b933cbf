#!/usr/bin/env python
"""cfg5 (TTA 6 views + WBF, 64 images): view-after-view eager vs TTAFusion.run on 1/2/3 streams vs the captured graph; L2 flushed."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heltondetection_b200 import synth, yolo, wbf
dev = torch.device("cuda")
B = 64
views, _ = synth.tta_heads(B, 640, 80, G=20, seed=1239)
vspec = [(r, flip, size) for (_, r, flip, size) in views]
devh = [[h.to(dev) for h in heads] for (heads, _, _, _) in views]
nbytes = sum(h.numel() * 4 for d in devh for h in d)
flush = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)


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
    print(f"{name:44s} {t * 1e3:8.1f} us  {nbytes / t / 1e6:7.1f} GB/s ({nbytes / t / 1e6 / 6536.7 * 100:.1f}% of measured peak)", flush=True)


for dense in (True, False):
    fusion = wbf.TTAFusion(vspec, (640, 640), 80, max_det=300, iou_thr=0.55, skip_box_thr=0.001)
    pps = [yolo.YoloPostprocessor(conf_thres=0.25, iou_thres=0.45, max_det=300, dense_read=dense, device=dev) for _ in views]

    def full():
        for v in range(len(views)):
            det, c, _ = pps[v](devh[v])
            fusion.map_back(v, det, c)
        return fusion.fuse()
    tag = "dense" if dense else "sparse"
    rep(f"[{tag}] eager, view after view", timed(full))
    for ns in (1, 2, 3):
        rep(f"[{tag}] run(), {ns} stream(s), eager", timed(lambda: fusion.run(pps, devh, ns)))
        replay, _ = fusion.graph(pps, devh, ns)
        rep(f"[{tag}] graph, {ns} stream(s)", timed(replay))
    rep(f"[{tag}] WBF alone", timed(lambda: fusion.fuse()))

#!/usr/bin/env python
"""cfg1 (one 640^2 image) at conf 0.001: which large-image NMS path is fastest for a single image (graph replay, L2 flushed)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heltondetection_b200 import synth, yolo, ops
dev = torch.device("cuda")
heads = [h.to(dev) for h in synth.yolo_heads(1, 640, 80, 20, 1234)[0]]
flush = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)


def timed(fn, iters=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        flush.zero_(); a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / iters


for conf in (0.25, 0.01, 0.001):
    for mode, name in ((0, "auto (clusters)"), (1, "one CTA per image"), (3, "light kernel")):
        ops.set_nms_mode(mode)
        pp = yolo.YoloPostprocessor(conf_thres=conf, iou_thres=0.45, max_det=300, dense_read=True, device=dev, one_call=False)
        rp, det, cnt, idx = pp.graph(heads)
        t = timed(rp)
        pd = yolo.YoloPostprocessor(conf_thres=conf, iou_thres=0.45, max_det=300, dense_read=True, device=dev, one_call=False)
        print(f"conf {conf:<6} {name:18s} {t * 1e3:7.1f} us   candidates {int(pp._buf.count[0])} kept {int(cnt[0])}", flush=True)
ops.set_nms_mode(0)

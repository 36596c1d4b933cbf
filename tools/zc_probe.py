#!/usr/bin/env python
"""zero-copy decode from pinned host heads (cfg2, 256 images): where the PCIe time goes.  conf 0.25 = the workload; conf 0.9999 = no
survivors, i.e. the objectness planes alone.  env HD_ZC_LINE=1: whole 128-byte lines around a survivor instead of 32-byte sectors."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heltondetection_b200 import synth, yolo
heads_cpu, _ = synth.yolo_heads(256, 640, 80, 20, 1235)
pinned = [h.contiguous().pin_memory() for h in heads_cpu]
dev = torch.device("cuda")
for conf in (0.25, 0.9999):
    pp = yolo.YoloPostprocessor(conf_thres=conf, iou_thres=0.45, dense_read=False, device=dev)
    for _ in range(3):
        det, cnt, idx = pp(pinned)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        det, cnt, idx = pp(pinned)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"HD_ZC_LINE={os.environ.get('HD_ZC_LINE', '0')} conf {conf}: {ms:7.3f} ms / 256 images  ({256 / ms * 1e3:8.0f} img/s)  kept/img {cnt.float().mean().item():.1f}", flush=True)

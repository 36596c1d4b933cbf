#!/usr/bin/env python
"""Phase clocks of block 0 of small_nms_kernel + event timings of the NMS entry at several batch sizes (developer tool)."""
import os, sys, ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heltondetection_b200 import synth, yolo, _lib, ops


def timeit(fn, iters=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def phases(tag):
    a = (C.c_longlong * 16)()
    _lib.check(_lib.lib().hd_debug_phases(2, a))
    v = list(a)
    names = ["load", "rank", "scatter", "mask", "resolve", "write", "zero tail"]
    print(tag, " ".join(f"{names[i]} {(v[i + 1] - v[i]) / 1.965e3:.2f}us" for i in range(7)), f"| total {(v[7] - v[0]) / 1.965e3:.2f} us")


for B in (1, 32, 256):
    heads, _ = synth.yolo_heads(B, 640, 80, 20, 1235)
    heads = [h.cuda() for h in heads]
    pp = yolo.YoloPostprocessor(conf_thres=0.25, iou_thres=0.45, dense_read=True, one_call=False)
    pp(heads)
    buf = pp._buf
    for mode in (0, 3):
        ops.set_nms_mode(mode)
        f = lambda: yolo._run_nms(buf, pp.iou_thres, pp.class_mode, pp.max_wh, pp.max_nms)
        g = torch.cuda.CUDAGraph()
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            f()
        print(f"B={B} nms mode {mode}: eager {timeit(f):.1f} us  graph {timeit(g.replay):.1f} us")
    ops.set_nms_mode(0)
    phases(f"B={B} small_nms block 0:")
    pp1 = yolo.YoloPostprocessor(conf_thres=0.25, iou_thres=0.45, dense_read=True)
    rp = pp1.graph(heads)[0]
    print(f"B={B} full step graph {timeit(rp):.1f} us")

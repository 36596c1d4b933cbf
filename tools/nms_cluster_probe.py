"""cfg4-style dense YOLO batch of B images: one CTA per image (mode 1) vs thread-block clusters (mode 0).  usage: python tools/nms_cluster_probe.py B"""
import sys, time, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heltondetection_b200 import synth, yolo, ops
B = int(sys.argv[1])
heads, _ = synth.yolo_heads(B, 1280, 10, 300, 1238, dense=True)
heads = [h.cuda() for h in heads]
for mode in (1, 0):
    ops.set_nms_mode(mode)
    pp = yolo.YoloPostprocessor(conf_thres=0.001, iou_thres=0.6)
    t0 = time.time()
    det, cnt, idx = pp(heads)
    torch.cuda.synchronize()
    print("mode", mode, "first call", time.time() - t0, "s; kept", cnt.tolist()[:4], flush=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): pp(heads)
    e1.record(); torch.cuda.synchronize()
    print("   per call", e0.elapsed_time(e1) / 5 * 1e3, "us", flush=True)

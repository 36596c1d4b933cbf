#!/usr/bin/env python
"""Fixed cost of a short timed region (the driver times 20 steps): 32-image shards on one GPU, depth 8, no peer.
20 steps as one 20-step cycle graph / 2x10 / 4x5 / step by step, against the steady state of 400 steps."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heltondetection_b200 import synth, yolo
dev = torch.device("cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
heads = [h.to(dev) for h in synth.yolo_heads(256, 640, 80, 20, 1235)[0]]
pool = [[h[j * B:(j + 1) * B] for h in heads] for j in range(256 // B)][:8]
depth = 8 if B < 128 else 4


def timed(pipe, n, reps=5):
    best = 1e9
    for _ in range(reps):
        pipe.fork(); pipe.run(0, n); pipe.join(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pipe.fork(); pipe.run(0, n); pipe.join(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e3


for cyc in (200, 20, 10, 5):
    pipe = yolo.PostprocessPipeline(pool, depth=depth, device=dev, cycle_graph=True, min_cycle=cyc, cycle_exact=True, conf_thres=0.25, iou_thres=0.45, max_det=300, dense_read=True)
    n = 400 if cyc == 200 else 20
    t = timed(pipe, n)
    print(f"B={B} cycle of {cyc:3d} steps: {n} steps in {t:8.1f} us = {t / n:6.2f} us/step", flush=True)
    if cyc == 20:
        pipe.cycle = None
        t = timed(pipe, 20)
        print(f"B={B} step by step      : 20 steps in {t:8.1f} us = {t / 20:6.2f} us/step", flush=True)
    del pipe
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); e1.record(); torch.cuda.synchronize()
print("empty region", e0.elapsed_time(e1) * 1e3, "us")

"""cfg4 dense batch through PostprocessPipeline at several depths (throughput form).  usage: python tools/cfg4_pipe.py [depths...]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heltondetection_b200 import synth, yolo
PEAK = 6536.7
heads = [h.cuda() for h in synth.yolo_heads(64, 1280, 10, 300, 1238, dense=True)[0]]
nbytes = sum(h.numel() * 4 for h in heads)
for depth in [int(a) for a in sys.argv[1:]] or [2, 3, 4, 6]:
    pl = yolo.PostprocessPipeline([heads], depth=depth, device=heads[0].device, cycle_graph=True, conf_thres=0.001, iou_thres=0.6, max_det=300, dense_read=True)
    pl.fork(); pl.run(0, pl.cycle_len); pl.join()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pl.fork(); pl.run(0, 4 * pl.cycle_len); pl.join(); e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (4 * pl.cycle_len)
    print(f"depth {depth}: {ms * 1e3:7.1f} us/step  {nbytes / ms / 1e6:7.1f} GB/s  {nbytes / ms / 1e6 / PEAK * 100:5.1f}% of measured peak  {64 / ms * 1e3:9.0f} img/s", flush=True)
    del pl

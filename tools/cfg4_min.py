"""cfg4 dense batch, a handful of post-process calls (target of ncu captures).  usage: python tools/cfg4_min.py [n_calls]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heltondetection_b200 import synth, yolo
heads = [h.cuda() for h in synth.yolo_heads(64, 1280, 10, 300, 1238, dense=True)[0]]
pp = yolo.YoloPostprocessor(conf_thres=0.001, iou_thres=0.6, dense_read=True, one_call=False)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    det, cnt, idx = pp(heads)
torch.cuda.synchronize()
print("kept/img", cnt.float().mean().item(), "candidates/img", pp._buf.count.float().mean().item())

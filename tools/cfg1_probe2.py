import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heltondetection_b200 import synth, yolo, ops, _lib
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda")
heads = [h.to(dev) for h in synth.yolo_heads(1, 640, 80, 20, 1234)[0]]
ops.set_nms_mode(0)
for one_call in (False, True):
    pp = yolo.YoloPostprocessor(conf_thres=0.001, iou_thres=0.45, max_det=300, dense_read=True, device=dev, one_call=one_call)
    for i in range(5):
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            pp(heads); torch.cuda.synchronize()
        print("one_call", one_call, "eager call", i, [(e.key[:34], round(e.device_time_total, 1)) for e in prof.key_averages() if "nms" in e.key])

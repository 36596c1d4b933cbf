import sys, os, ctypes as C, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heltondetection_b200 import synth, rpn, yolo, _lib
def phases(tag, which):
    a = (C.c_longlong * 16)()
    _lib.check(_lib.lib().hd_debug_phases(which, a))
    v = list(a)
    names = {0: "start", 1: "select done", 2: "compaction done", 3: "sort done", 4: "boxes materialised", 8: "grid built", 9: "16 chunks", 5: "nms done", 6: "outputs"}
    order = [0, 1, 2, 3, 4, 8, 9, 5, 6]
    prev = None
    print(tag)
    for i in order:
        if v[i] == 0: continue
        if prev is not None and v[i] >= prev: print(f"   {names[i]:22s} +{(v[i]-prev)/1.9e3:8.1f} us")
        prev = v[i]
obj, dlt, bases, _ = synth.rpn_heads(1, 832, G=20, seed=1237)
pr = rpn.RpnProposals(bases, (4, 8, 16, 32), (832, 832), n_pre_nms=12000, n_post_nms=2000, min_size=16)
for _ in range(2): pr([o.cuda() for o in obj], [d.cuda() for d in dlt])
phases("rpn_select_nms_kernel (1 image, n_pre 12000)", 1)
heads, _ = synth.yolo_heads(1, 1280, 10, 300, 1238, dense=True)
pp = yolo.YoloPostprocessor(conf_thres=0.001, iou_thres=0.6)
for _ in range(2): pp([h.cuda() for h in heads])
phases("sort_nms_kernel (cfg4, 1 image, n~6800)", 0)

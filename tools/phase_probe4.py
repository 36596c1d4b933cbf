import sys, os, ctypes as C, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heltondetection_b200 import synth, yolo, _lib, ops
def phases(tag, which):
    a = (C.c_longlong * 16)()
    _lib.check(_lib.lib().hd_debug_phases(which, a))
    v = list(a)
    names = {0: "start", 2: "keys built", 3: "sort done", 4: "boxes materialised", 8: "grid built", 9: "16 chunks", 5: "nms done"}
    order = [0, 2, 3, 4, 8, 9, 5]
    prev = None
    print(tag)
    for i in order:
        if v[i] == 0: continue
        if prev is not None and v[i] >= prev: print(f"   {names[i]:22s} +{(v[i]-prev)/1.965e3:8.1f} us")
        prev = v[i]
ops.set_nms_mode(4)
for B in (1, 64):
    heads, _ = synth.yolo_heads(B, 1280, 10, 300, 1238, dense=True)
    pp = yolo.YoloPostprocessor(conf_thres=0.001, iou_thres=0.6, one_call=False)
    for _ in range(3): pp([h.cuda() for h in heads])
    torch.cuda.synchronize()
    print("B", B, "candidates", pp._buf.count[:4].tolist())
    phases(f"sort_nms_kernel<1024> block 0 (cfg4, B={B})", 0)

"""Dense YOLO batches of B images: one CTA per image (mode 1, bucket sort) vs thread-block clusters (mode 0/2); graph replays, L2 flushed.
usage: python tools/nms_cluster_probe.py B [B ...]"""
import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heltondetection_b200 import synth, yolo, ops
flush = torch.empty((256 << 20,), dtype=torch.uint8, device="cuda")


def timed(fn, iters=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        flush.zero_(); a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / iters


for B in [int(a) for a in sys.argv[1:]] or [1, 4, 16, 32]:
    for tag, (img, nc, G, seed, dense, conf, iou) in (("cfg4-like 1280^2 nc=10", (1280, 10, 300, 1238, True, 0.001, 0.6)),
                                                       ("cfg1-like 640^2 nc=80 conf .001", (640, 80, 20, 1234, False, 0.001, 0.45))):
        heads = [h.cuda() for h in synth.yolo_heads(B, img, nc, G, seed, dense=dense)[0]]
        res = {}
        for mode in (1, 0):
            ops.set_nms_mode(mode)
            pp = yolo.YoloPostprocessor(conf_thres=conf, iou_thres=iou, dense_read=True, one_call=False)
            rp, det, cnt, idx = pp.graph(heads)
            res[mode] = (timed(rp), idx.clone(), int(pp._buf.count.float().mean()))
        ops.set_nms_mode(0)
        print(f"B={B:3d} {tag:34s} candidates/img {res[1][2]:5d}: one CTA per image {res[1][0] * 1e3:7.1f} us   auto (clusters) {res[0][0] * 1e3:7.1f} us   same keeps {bool(torch.equal(res[0][1], res[1][1]))}", flush=True)
        del heads

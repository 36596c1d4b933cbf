#!/usr/bin/env python
"""Stage-by-stage device timings of the five BASELINE.json configs (developer tool, not the bench contract).
usage: python tools/perf_probe.py [cfg2 cfg3 cfg4 cfg5 zerocopy]"""
import os
import sys
import json
import time
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heltondetection_b200 import synth, yolo, rpn, roi, wbf, _lib  # noqa: E402

PEAK = 6536.7


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def report(name, ms, nbytes=None, imgs=None):
    s = f"{name:44s} {ms * 1e3:10.1f} us"
    if nbytes:
        s += f"  {nbytes / ms / 1e6:8.1f} GB/s ({nbytes / ms / 1e6 / PEAK * 100:5.1f}% of measured peak)"
    if imgs:
        s += f"  {imgs / ms * 1e3:10.0f} img/s"
    print(s, flush=True)


def yolo_cfg(tag, B, img, nc, G, seed, conf, iou, dense_scene):
    heads_cpu, _ = synth.yolo_heads(B, img, nc, G, seed, dense=dense_scene)
    heads = [h.cuda() for h in heads_cpu]
    nbytes = sum(h.numel() * 4 for h in heads)
    for mode in (True, False):
        pp = yolo.YoloPostprocessor(conf_thres=conf, iou_thres=iou, dense_read=mode)
        arr, keep, Bn, A, ncc, total = yolo._levels(heads, pp.anchors, pp.strides)
        buf = pp.buffers(Bn, total, heads[0].device)
        L = _lib.lib()

        def dec():
            _lib.check(L.hd_yolo_decode_filter(arr, len(keep), Bn, A, ncc, pp.conf_thres, pp.flags, _lib.ptr(buf.box), _lib.ptr(buf.score),
                                               _lib.ptr(buf.cls), _lib.ptr(buf.anchor), _lib.ptr(buf.count), buf.cap, _lib.stream()))
        t = timeit(dec)
        report(f"{tag} decode_filter {'dense' if mode else 'sparse'}", t, nbytes, B)
        t2 = timeit(lambda: yolo._run_nms(buf, pp.iou_thres, pp.class_mode, pp.max_wh, pp.max_nms))
        report(f"{tag} sort_nms", t2, None, B)
        cnt = buf.count.float()
        print(f"    candidates/img mean {cnt.mean().item():.0f} max {cnt.max().item():.0f}; kept mean {buf.out_count.float().mean().item():.0f}")
        t3 = timeit(lambda: pp(heads))
        report(f"{tag} full postprocess ({'dense' if mode else 'sparse'})", t3, nbytes, B)
        rp = pp.graph(heads)[0]
        t4 = timeit(rp, iters=50)
        report(f"{tag} full postprocess, CUDA graph ({'dense' if mode else 'sparse'})", t4, nbytes, B)
        pu = yolo.YoloPostprocessor(conf_thres=conf, iou_thres=iou, dense_read=mode, one_call=False)
        t5 = timeit(pu.graph(heads)[0], iters=50)
        report(f"{tag} two-kernel path, CUDA graph ({'dense' if mode else 'sparse'})", t5, nbytes, B)
    t = timeit(lambda: yolo.decode_box(heads), iters=5)
    report(f"{tag} decode_box (dense pred out)", t, 2 * nbytes, B)
    return heads_cpu, heads


def cfg1():
    """the reference's own CPU-runnable case: one image"""
    yolo_cfg("cfg1", 1, 640, 80, 20, 1234, 0.25, 0.45, False)


def cfg2():
    yolo_cfg("cfg2", 256, 640, 80, 20, 1235, 0.25, 0.45, False)


def cfg2_half():
    """8f-4: the same cfg2 batch with fp16 / bf16 heads (widened to fp32 in the kernel): half the bytes"""
    heads_cpu, _ = synth.yolo_heads(256, 640, 80, 20, 1235)
    for dt in (torch.float16, torch.bfloat16):
        heads = [h.to(dt).cuda() for h in heads_cpu]
        nbytes = sum(h.numel() * 2 for h in heads)
        for mode in (True, False):
            pp = yolo.YoloPostprocessor(conf_thres=0.25, iou_thres=0.45, dense_read=mode)
            rp, det, cnt, idx = pp.graph(heads)
            t = timeit(rp, iters=50)
            report(f"cfg2 {str(dt)[6:]} heads, full postprocess, graph ({'dense' if mode else 'sparse'})", t, nbytes, 256)
        del heads
    heads = [h.cuda().contiguous(memory_format=torch.channels_last) for h in heads_cpu]
    nbytes = sum(h.numel() * 4 for h in heads)
    for mode in (True, False):
        pp = yolo.YoloPostprocessor(conf_thres=0.25, iou_thres=0.45, dense_read=mode)
        rp, det, cnt, idx = pp.graph(heads)
        t = timeit(rp, iters=50)
        report(f"cfg2 fp32 channels_last (NHWC) heads, full postprocess, graph ({'dense' if mode else 'sparse'})", t, nbytes, 256)


def cfg4():
    yolo_cfg("cfg4", 64, 1280, 10, 300, 1238, 0.001, 0.6, True)


def cfg3():
    B, img = 16, 832
    obj, dlt, bases, _ = synth.rpn_heads(B, img, G=20, seed=1237)
    feats = [f.cuda() for f in synth.fpn_features(B, img, 256, 1237)]
    obj, dlt = [o.cuda() for o in obj], [d.cuda() for d in dlt]
    pr = rpn.RpnProposals(bases, (4, 8, 16, 32), (img, img), n_pre_nms=12000, n_post_nms=2000, min_size=16)
    rpn_bytes = sum(o.numel() * 4 for o in obj) + sum(d.numel() * 4 for d in dlt)
    t = timeit(lambda: pr.decode(obj, dlt), iters=10)
    report("cfg3 rpn decode (incl. torch.empty)", t, rpn_bytes, B)
    t = timeit(lambda: pr(obj, dlt), iters=5)
    report("cfg3 rpn decode+select+nms", t, rpn_bytes, B)
    rois, cnt, sc, idx = pr(obj, dlt)
    print("    proposals kept/img:", cnt.tolist())
    pl = rpn.RpnProposalsPerLevel(bases, (4, 8, 16, 32), (img, img), nms_thresh=0.7, pre_nms_top_n=1000, post_nms_top_n=1000)
    t = timeit(lambda: pl(obj, dlt), iters=5)
    report("cfg3 rpn per-level variant (1000/level, torchvision semantics)", t, rpn_bytes, B)
    scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
    fbytes = sum(f.numel() * 4 for f in feats)
    obytes = rois.shape[0] * 256 * 49 * 4
    nhwc = [f.contiguous(memory_format=torch.channels_last) for f in feats]
    for sr in (2, 0):
        t = timeit(lambda: roi.multilevel_roi_align(nhwc, rois, 7, scales, sr, False), iters=5)
        report(f"cfg3 multilevel roi_align NHWC-in sr={sr}", t, fbytes + obytes, B)
    t = timeit(lambda: roi.multilevel_roi_align(feats, rois, 7, scales, 2, False), iters=5)
    report("cfg3 multilevel roi_align NCHW-in (+layout pass)", t, fbytes + obytes, B)
    t = timeit(lambda: roi.multilevel_roi_align(nhwc, rois, 7, scales, 2, False, op="pool"), iters=5)
    report("cfg3 multilevel roi_pool NHWC-in", t, fbytes + obytes, B)
    t = timeit(lambda: roi.multilevel_roi_align(nhwc[:1], rois, 7, scales[:1], 2, False), iters=3)
    report("cfg3 single-level P2 roi_align NHWC sr=2", t, feats[0].numel() * 4 + obytes, B)
    t = timeit(lambda: roi.multilevel_roi_align(nhwc[:1], rois, 7, scales[:1], 0, False), iters=2)
    report("cfg3 single-level P2 roi_align NHWC sr=0", t, feats[0].numel() * 4 + obytes, B)
    out = torch.empty((B, 208, 208, 256), device="cuda")
    t = timeit(lambda: _lib.check(_lib.lib().hd_nchw_to_nhwc(_lib.ptr(feats[0]), _lib.ptr(out), B, 256, 208, 208, _lib.stream())), iters=5)
    report("cfg3 nchw->nhwc level0", t, 2 * feats[0].numel() * 4)
    t = timeit(lambda: (pr(obj, dlt), roi.multilevel_roi_align(nhwc, rois, 7, scales, 2, False)), iters=5)
    report("cfg3 RPN + multilevel RoIAlign (NHWC in), per batch", t, rpn_bytes + fbytes + obytes, B)
    try:
        import torchvision
        small = rois[:4000]
        t = timeit(lambda: torchvision.ops.roi_align(feats[1], small, 7, 1 / 8, 2, False), iters=3)
        report("   (torchvision CUDA roi_align, 4000 rois, lvl1)", t, None)
        t = timeit(lambda: roi.roi_align(nhwc[1], small, 7, 1 / 8, 2, False), iters=3)
        report("   (hd_b200 NHWC roi_align, 4000 rois, lvl1)", t, None)
        bx = rois[:12000, 1:].contiguous(); ss = torch.rand(12000, device="cuda")
        t = timeit(lambda: torchvision.ops.nms(bx, ss, 0.7), iters=3)
        report("   (torchvision CUDA nms n=12000)", t, None)
    except Exception as e:
        print("   torchvision cuda comparison skipped:", e)


def roihead():
    """FasterRCNN final stage on the cfg3 proposals: 16 x 2000 RoIs x 81 classes -> softmax/decode/filter -> class-aware NMS -> top 100"""
    from heltondetection_b200 import roi_head
    B, R, Cn = 16, 2000, 81
    g = torch.Generator().manual_seed(1240)
    lg = (torch.randn(B * R, Cn, generator=g) * 3).cuda()
    rg = (torch.randn(B * R, Cn * 4, generator=g) * 0.5).cuda()
    xy = torch.rand(B * R, 2, generator=g) * 700
    rois = torch.cat((torch.arange(B).repeat_interleave(R)[:, None].float(), xy, xy + torch.rand(B * R, 2, generator=g) * 120 + 8), 1).cuda()
    pp = roi_head.RoIHeadPostprocessor((832, 832), 0.05, 0.5, 100)
    nbytes = lg.numel() * 4 + rg.numel() * 4 + rois.numel() * 4
    t = timeit(lambda: pp(lg, rg, rois, None, B=B), iters=10)
    det, idx, cnt = pp(lg, rg, rois, None, B=B)
    report("roi-head post-process 16x2000x81 (decode+filter+NMS+top100)", t, nbytes, B)
    print("    detections/img:", cnt.tolist()[:8])
    t = timeit(lambda: roi_head.scale_coords((832, 832), det, [(1080, 1920)] * B, cnt, xywh=True), iters=10)
    report("scale_coords + COCO boxes 16x100", t, None)


def cfg5():
    B, img, nc = 64, 640, 80
    views, _ = synth.tta_heads(B, img, nc, G=20, seed=1239)
    vspec = [(r, flip, size) for (_, r, flip, size) in views]
    dev = [[h.cuda() for h in heads] for (heads, _, _, _) in views]
    fusion = wbf.TTAFusion(vspec, (img, img), nc, max_det=300, iou_thr=0.55, skip_box_thr=0.001)
    pps = [yolo.YoloPostprocessor(conf_thres=0.25, iou_thres=0.45) for _ in views]

    def full():
        for v in range(len(views)):
            det, cnt, _ = pps[v](dev[v])
            fusion.map_back(v, det, cnt)
        return fusion.fuse()
    nbytes = sum(h.numel() * 4 for d in dev for h in d)
    t = timeit(full, iters=10)
    report("cfg5 6x(decode+nms+mapback) + wbf (sparse)", t, nbytes, B)
    full()
    t = timeit(lambda: fusion.fuse(), iters=10)
    report("cfg5 wbf alone", t, None, B)
    print("    fused/img mean", fusion.wbf.oc.float().mean().item(), "inputs/img", fusion.cnt.sum(1).float().mean().item())


def zerocopy():
    """sparse decode reading pinned HOST memory directly over PCIe (UVA zero-copy)."""
    B = 256
    heads_cpu, _ = synth.yolo_heads(B, 640, 80, 20, 1235)
    pinned = [h.pin_memory() for h in heads_cpu]
    nbytes = sum(h.numel() * 4 for h in heads_cpu)
    pp = yolo.YoloPostprocessor(conf_thres=0.25, iou_thres=0.45, dense_read=False)
    dev = [h.cuda() for h in heads_cpu]
    det_ref, cnt_ref, idx_ref = pp(dev)
    det_ref, cnt_ref, idx_ref = det_ref.clone(), cnt_ref.clone(), idx_ref.clone()
    ppz = yolo.YoloPostprocessor(conf_thres=0.25, iou_thres=0.45, dense_read=False, device="cuda:0")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    det, cnt, idx = ppz(pinned)
    torch.cuda.synchronize()
    print("zero-copy first call", (time.perf_counter() - t0) * 1e3, "ms")
    print("zero-copy parity: counts equal", bool(torch.equal(cnt, cnt_ref)), "idx equal", bool(torch.equal(idx, idx_ref)))
    t = timeit(lambda: ppz(pinned), iters=5)
    report("zero-copy sparse postprocess from pinned host", t, nbytes, B)
    stage = [torch.empty_like(h, device="cuda") for h in heads_cpu]

    def copy_then():
        for s, p in zip(stage, pinned):
            s.copy_(p, non_blocking=True)
        pp(stage)
    t = timeit(copy_then, iters=5)
    report("H2D copy + sparse postprocess", t, nbytes, B)


def cpu():
    """The reference's CPU path (oracle = torch CPU + torchvision CPU ops + numpy WBF) on bounded samples of every config."""
    import oracle
    import numpy as np
    torch.set_num_threads(os.cpu_count() or 1)
    nt = torch.get_num_threads()

    def best(fn, reps=3):
        fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        return min(ts)

    def line(name, t, n_img):
        print(f"CPU[{nt} threads] {name:58s} {t * 1e3:9.1f} ms for {n_img} img -> {n_img / t:9.1f} img/s", flush=True)

    h, _ = synth.yolo_heads(16, 640, 80, 20, 1235)
    line("cfg2 decode_box + non_max_suppression (conf .25)", best(lambda: oracle.yolo.non_max_suppression(oracle.yolo.decode_box(h), 0.25, 0.45)), 16)
    line("cfg1 same, batch 1", best(lambda: oracle.yolo.non_max_suppression(oracle.yolo.decode_box([x[:1] for x in h]), 0.25, 0.45)), 1)
    h4, _ = synth.yolo_heads(2, 1280, 10, 300, 1238, dense=True)
    line("cfg4 1280^2 nc=10 dense, conf .001 iou .6", best(lambda: oracle.yolo.non_max_suppression(oracle.yolo.decode_box(h4), 0.001, 0.6)), 2)
    obj, dlt, bases, _ = synth.rpn_heads(2, 832, G=20, seed=1237)
    props = oracle.rpn.rpn_proposals(obj, dlt, bases, (4, 8, 16, 32), (832, 832), n_pre_nms=12000, n_post_nms=2000, min_size=16)
    line("cfg3 RPN decode + top-12000 + nms(0.7) + top-2000", best(lambda: oracle.rpn.rpn_proposals(obj, dlt, bases, (4, 8, 16, 32), (832, 832), n_pre_nms=12000, n_post_nms=2000, min_size=16), 2), 2)
    feats = synth.fpn_features(1, 832, 256, 1237)
    r = torch.cat((torch.zeros(len(props[0][0]), 1), props[0][0]), 1)
    line("cfg3 multi-level RoIAlign 2000x256x7x7 (sr=2)", best(lambda: oracle.roi.multilevel_roi_align(feats, r, 7, [1 / 4, 1 / 8, 1 / 16, 1 / 32], 2, False), 2), 1)
    line("cfg3 multi-level RoIPool 2000x256x7x7", best(lambda: oracle.roi.multilevel_roi_align(feats, r, 7, [1 / 4, 1 / 8, 1 / 16, 1 / 32], 2, False, op="pool"), 2), 1)
    views, _ = synth.tta_heads(2, 640, 80, G=20, seed=1239)

    def tta():
        for b in range(2):
            bl, sl, ll = [], [], []
            for (hh, rr, ff, ss) in views:
                d = oracle.yolo.non_max_suppression(oracle.yolo.decode_box([x[b:b + 1] for x in hh]), 0.25, 0.45)[0]
                bb, sc, lb = oracle.tta.map_back(d, rr, ff, float(ss), 640.0, 640.0)
                bl.append(bb.numpy()); sl.append(sc.numpy()); ll.append(lb.numpy())
            oracle.wbf.weighted_boxes_fusion(bl, sl, ll, None, 0.55, 0.001)
    line("cfg5 TTA x6 (decode+NMS) + map-back + WBF", best(tta, 2), 2)


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg1", "cfg2", "cfg2_half", "cfg3", "roihead", "cfg4", "cfg5", "zerocopy", "cpu"]
    torch.cuda.set_device(0)
    for w in which:
        print(f"==== {w}", flush=True)
        globals()[w]()

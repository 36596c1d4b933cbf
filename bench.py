#!/usr/bin/env python
"""bench.py -- post-process images/sec of the YOLOv5 decode+filter+NMS hot path (BASELINE.json configs[1]:
YOLOv5s 640x640, batch 256, 80 classes, 25 200 anchors, image-sharded over N GPUs).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scaling strong|weak] [--depth D] [--no-configs]

One "step" = one pass of the hot path over one global batch of 256 synthetic images.  With N ranks the batch is SHARDED
(strong scaling, BASELINE configs[1] / SURVEY.md 8e: 256 images -> 256/N per rank, 32 at N=8): every rank runs the fused
decode+filter kernel and the NMS kernels on its slice, and the NMS kernels store the kept rows into every rank's gather buffer
over NVLink (the all-gather fused into the kernel epilogue; one cross-GPU barrier per step).  Steps are software-pipelined
D deep over CUDA streams (yolo.PostprocessPipeline): the NMS of step k overlaps the HBM-bound decode of step k+1.  Every rank
rotates through a pool of N distinct shard inputs (2.19 GB per rank at every N), so the cache behaviour does not change with N.
`--scaling weak` keeps the round-1 behaviour (256 images per rank).  Rank 0 prints ONE JSON line; at N=1 it also carries a
`configs` object with device-timed lines for BASELINE configs 1, 3, 4, 5 (L2 flushed between iterations).
"""
import argparse
import json
import math
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IMG, NC, G, BATCH = 640, 80, 20, 256
CONF, IOU, MAX_DET = 0.25, 0.45, 300
BYTES_PER_IMG = 25200 * 85 * 4  # 8 568 000 B: every head element once
WORKLOAD = "cfg2: YOLOv5s 640x640 global batch 256 nc=80 25200 anchors decode+conf-filter+class-aware NMS"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="dense", choices=["dense", "sparse"],
                    help="dense: every head byte is read (roofline-honest headline); sparse: objectness-tile skip")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: the global batch is sharded over the ranks (the contracted config); weak: --batch images per rank")
    ap.add_argument("--depth", type=int, default=0, help="software-pipeline depth (CUDA streams); 0 = auto: 4 for shards of >= 128 images, "
                    "8 below (measured on B200: a 32-image shard runs 45.9 / 41.0 / 38.8 / 38.5 us per step at depth 3 / 4 / 6 / 8)")
    ap.add_argument("--batch", type=int, default=BATCH, help="global batch (strong) / per-rank batch (weak)")
    ap.add_argument("--e2e-steps", type=int, default=24)
    ap.add_argument("--cpu-sample", type=int, default=32)
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg1/3/4/5 lines (they are only produced at N=1)")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.recording = False   # the thread starts during warm-up (first NVML calls are slow); only the timed region is kept
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                if self.recording:
                    self.samples.append(mhz)
                    for bit, name in names.items():
                        if r & bit:
                            self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=1.0)
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ CPU reference legs
def _best(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


def cpu_reference_rate(heads_cpu, n_img, repeats=3):
    """The oracle (torch CPU + torchvision CPU ops, the reference's CPU path) on n_img images."""
    import oracle
    sub = [h[:n_img] for h in heads_cpu]
    return n_img / _best(lambda: oracle.yolo.non_max_suppression(oracle.yolo.decode_box(sub), CONF, IOU, max_det=MAX_DET), repeats)


def run_reference(args, rank):
    import torch
    if rank != 0:
        return
    from heltondetection_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    n = min(args.cpu_sample, args.batch)
    heads, _ = synth.yolo_heads(n, IMG, NC, G, 1235)
    import oracle
    for _ in range(3):
        oracle.yolo.non_max_suppression(oracle.yolo.decode_box(heads), CONF, IOU, max_det=MAX_DET)
    steps = min(args.steps, 20)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle.yolo.non_max_suppression(oracle.yolo.decode_box(heads), CONF, IOU, max_det=MAX_DET)
    dt = time.perf_counter() - t0
    v = n * steps / dt
    sample = f"{n} images/step x {steps} steps of the same synthetic workload (oracle = torch CPU + torchvision CPU ops)"
    print(json.dumps({
        "impl": "reference", "metric": "post-process images/sec", "value": v, "unit": "img/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": 3, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": v, "unit": "img/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# ------------------------------------------------------------------------------------------------ cfg1/3/4/5 lines (N=1)
class Flusher:
    """writes a buffer larger than the 126 MB L2 between timed iterations"""

    def __init__(self, dev, mb=256):
        import torch
        self.buf = torch.empty((mb << 20,), dtype=torch.uint8, device=dev)

    def __call__(self):
        self.buf.zero_()


def time_flushed(fn, flush, iters=20, warm=3):
    """average device time of fn() over `iters` iterations, L2 flushed before each one (flush not timed) -> ms"""
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        flush()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs) / iters


def roof(nbytes, ms, peak):
    ach = nbytes / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
            "algorithmic_bytes_per_step": nbytes}


def extra_configs(dev, peak, cpu_cores):
    """Device-timed lines for BASELINE configs 1, 3, 4, 5 (one GPU), each with its roofline fraction (algorithmic bytes of SURVEY.md 8d
    over the measured step), a bounded CPU sample of the same workload through the oracle, and -- for nms / roi_align -- the
    torchvision CUDA kernels on the same inputs on the same box."""
    import numpy as np
    import torch
    import oracle
    from heltondetection_b200 import synth, yolo, rpn, roi, wbf
    out = {}
    flush = Flusher(dev)
    torch.set_num_threads(cpu_cores)

    def yolo_line(B, img, nc, Gn, seed, conf, iou, dense_scene, cpu_n, iters):
        heads_cpu, _ = synth.yolo_heads(B, img, nc, Gn, seed, dense=dense_scene)
        heads = [h.to(dev) for h in heads_cpu]
        nbytes = sum(h.numel() * 4 for h in heads)
        pp = yolo.YoloPostprocessor(conf_thres=conf, iou_thres=iou, max_det=MAX_DET, dense_read=True, device=dev)
        rp, det, cnt, idx = pp.graph(heads)
        ms = time_flushed(rp, flush, iters)
        sub = [h[:cpu_n] for h in heads_cpu]
        cpu_t = _best(lambda: oracle.yolo.non_max_suppression(oracle.yolo.decode_box(sub), conf, iou, max_det=MAX_DET), 2)
        ref = oracle.yolo.non_max_suppression(oracle.yolo.decode_box(sub), conf, iou, max_det=MAX_DET, return_index=True)[1]
        got = [idx[b, :int(cnt[b])].cpu() for b in range(cpu_n)]
        pipe_ms = None
        if B >= 16:   # throughput form: steps pipelined over 3 streams (NMS of step k under the decode of step k+1); inputs > L2
            pl = yolo.PostprocessPipeline([heads], depth=3, device=dev, cycle_graph=True, conf_thres=conf, iou_thres=iou, max_det=MAX_DET, dense_read=True)
            pl.fork(); pl.run(0, pl.cycle_len); pl.join()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); pl.fork(); pl.run(0, 2 * pl.cycle_len); pl.join(); e1.record()
            torch.cuda.synchronize()
            pipe_ms = e0.elapsed_time(e1) / (2 * pl.cycle_len)
            del pl
        line = {"ms": ms, "img_s": B / ms * 1e3, "roofline": roof(nbytes, ms, peak), "batch": B,
                "candidates_per_img": None, "kept_per_img": float(cnt.float().mean()),
                "keep_indices_match_oracle": bool(all(torch.equal(a, b) for a, b in zip(got, ref))),
                "cpu_baseline": {"value": cpu_n / cpu_t, "unit": "img/s", "cores": cpu_cores, "kind": "port", "sample": f"{cpu_n} images of the batch, best of 2"}}
        if pipe_ms is not None:
            line["pipelined"] = {"ms": pipe_ms, "img_s": B / pipe_ms * 1e3, "roofline": roof(nbytes, pipe_ms, peak),
                                 "note": "steps software-pipelined 3 deep over CUDA streams (the form bench.py's headline uses); inputs larger than L2, no flush"}
        del heads
        return line

    # cfg1: the reference's own CPU-runnable case, one image; both thresholds of SURVEY.md 8d
    out["cfg1_conf0.25"] = yolo_line(1, 640, 80, 20, 1234, 0.25, 0.45, False, 1, 50)
    out["cfg1_conf0.001"] = yolo_line(1, 640, 80, 20, 1234, 0.001, 0.45, False, 1, 50)
    out["cfg1_conf0.25"]["workload"] = out["cfg1_conf0.001"]["workload"] = "cfg1: YOLOv5s 640x640 batch 1 nc=80 decode+filter+NMS (latency-bound: one image is 8.6 MB)"
    # cfg4: dense VisDrone-style scenes
    out["cfg4"] = yolo_line(64, 1280, 10, 300, 1238, 0.001, 0.6, True, 2, 20)
    out["cfg4"]["workload"] = "cfg4: YOLOv5l 1280x1280 batch 64 nc=10 dense scenes conf 0.001 iou 0.6 (NMS-bound, ~6.8k candidates/img)"

    # cfg3: RPN proposals + multi-level RoIAlign
    B, img = 16, 832
    obj_c, dlt_c, bases, _ = synth.rpn_heads(B, img, G=20, seed=1237)
    feats_c = synth.fpn_features(B, img, 256, 1237)
    obj, dlt = [o.to(dev) for o in obj_c], [d.to(dev) for d in dlt_c]
    nhwc = [f.to(dev).contiguous(memory_format=torch.channels_last) for f in feats_c]
    scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
    pr = rpn.RpnProposals(bases, (4, 8, 16, 32), (img, img), n_pre_nms=12000, n_post_nms=2000, min_size=16)
    rois, cnt, sc, idx = pr(obj, dlt)
    rpn_bytes = sum(o.numel() * 4 for o in obj) + sum(d.numel() * 4 for d in dlt)
    fbytes = sum(f.numel() * 4 for f in nhwc)
    obytes = rois.shape[0] * 256 * 49 * 4
    t_rpn = time_flushed(lambda: pr(obj, dlt), flush, 10)
    t_roi = time_flushed(lambda: roi.multilevel_roi_align(nhwc, rois, 7, scales, 2, False), flush, 10)
    t_roi0 = time_flushed(lambda: roi.multilevel_roi_align(nhwc, rois, 7, scales, 0, False), flush, 5)
    t_pool = time_flushed(lambda: roi.multilevel_roi_align(nhwc, rois, 7, scales, 2, False, op="pool"), flush, 5)
    t_both = time_flushed(lambda: (pr(obj, dlt), roi.multilevel_roi_align(nhwc, rois, 7, scales, 2, False)), flush, 10)
    # throughput form: batch k+1's (latency-bound) RPN runs under batch k's (bandwidth-bound) RoIAlign, two streams
    pr2 = rpn.RpnProposals(bases, (4, 8, 16, 32), (img, img), n_pre_nms=12000, n_post_nms=2000, min_size=16)
    st2, prs = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)], [pr, pr2]

    def cfg3_steps(n):
        ev = torch.cuda.Event(); ev.record()
        for s in st2:
            s.wait_event(ev)
        for k in range(n):
            with torch.cuda.stream(st2[k & 1]):
                r_, _, _, _ = prs[k & 1](obj, dlt)
                roi.multilevel_roi_align(nhwc, r_, 7, scales, 2, False)
        for s in st2:
            torch.cuda.current_stream().wait_stream(s)
    cfg3_steps(4)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); cfg3_steps(12); e1.record()
    torch.cuda.synchronize()
    t_pipe = e0.elapsed_time(e1) / 12
    cpu_rpn = _best(lambda: oracle.rpn.rpn_proposals([o[:1] for o in obj_c], [d[:1] for d in dlt_c], bases, (4, 8, 16, 32), (img, img),
                                                     n_pre_nms=12000, n_post_nms=2000, min_size=16), 1)
    r0 = rois[:2000].cpu()
    cpu_roi = _best(lambda: oracle.roi.multilevel_roi_align([f[:1] for f in feats_c], r0, 7, scales, 2, False), 1)
    line = {"workload": "cfg3: FasterRCNN-PAFPN 832x832 batch 16: RPN decode + top-12000 + NMS 0.7 + top-2000, then multi-level RoIAlign 7x7x256 (sr=2, NHWC features)",
            "ms": t_both, "img_s": B / t_both * 1e3, "roofline": roof(rpn_bytes + fbytes + obytes + rois.numel() * 4, t_both, peak),
            "rpn_ms": t_rpn, "roi_align_ms": t_roi, "roi_align_roofline": roof(fbytes + obytes, t_roi, peak),
            "roi_align_adaptive_sr0_ms": t_roi0, "roi_pool_ms": t_pool, "proposals_per_img": int(cnt.float().mean()),
            "pipelined": {"ms": t_pipe, "img_s": B / t_pipe * 1e3, "roofline": roof(rpn_bytes + fbytes + obytes + rois.numel() * 4, t_pipe, peak),
                          "note": "batch k+1's RPN on a second stream under batch k's RoIAlign; features (0.94 GB) larger than L2, no flush"},
            "cpu_baseline": {"value": 1.0 / (cpu_rpn + cpu_roi), "unit": "img/s", "cores": cpu_cores, "kind": "port",
                             "sample": "1 image: oracle RPN proposals + torchvision CPU multi-level roi_align, best of 1 after 1 warm-up"}}
    try:   # the torchvision sm_100 kernels on the same inputs, same box (comparison only; never on the product path)
        import torchvision
        lv1 = rois[(roi.level_map(rois[:, 1:]) == 1)][:4000].contiguous()
        if lv1.shape[0] < 64:
            lv1 = rois[:4000]
        nchw1 = nhwc[1].contiguous()
        t_tv = time_flushed(lambda: torchvision.ops.roi_align(nchw1, lv1, 7, 1 / 8, 2, False), flush, 5)
        t_hd = time_flushed(lambda: roi.roi_align(nhwc[1], lv1, 7, 1 / 8, 2, False), flush, 5)
        bx = rois[:12000, 1:].contiguous()
        ss = torch.rand(12000, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
        from heltondetection_b200 import ops
        t_tvn = time_flushed(lambda: torchvision.ops.nms(bx, ss, 0.7), flush, 5)
        t_hdn = time_flushed(lambda: ops.nms(bx, ss, 0.7), flush, 5)
        same = bool(torch.equal(torchvision.ops.nms(bx, ss, 0.7), ops.nms(bx, ss, 0.7)))
        line["torchvision_cuda_same_box"] = {"roi_align_ms": {"torchvision": t_tv, "hd_b200": t_hd, "rois": int(lv1.shape[0])},
                                             "nms_n12000_ms": {"torchvision": t_tvn, "hd_b200": t_hdn, "keep_equal": same}}
    except Exception as e:  # noqa: BLE001
        line["torchvision_cuda_same_box"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    out["cfg3"] = line
    del obj, dlt, nhwc

    # cfg5: TTA (3 scales x {id, hflip}) + WBF
    B = 64
    views, _ = synth.tta_heads(B, 640, 80, G=20, seed=1239)
    vspec = [(r, flip, size) for (_, r, flip, size) in views]
    devh = [[h.to(dev) for h in heads] for (heads, _, _, _) in views]
    fusion = wbf.TTAFusion(vspec, (640, 640), 80, max_det=MAX_DET, iou_thr=0.55, skip_box_thr=0.001)
    pps = [yolo.YoloPostprocessor(conf_thres=CONF, iou_thres=IOU, max_det=MAX_DET, dense_read=True, device=dev) for _ in views]

    def full():
        for v in range(len(views)):
            det, c, _ = pps[v](devh[v])
            fusion.map_back(v, det, c)
        return fusion.fuse()
    nbytes = sum(h.numel() * 4 for d in devh for h in d)
    t_eager = time_flushed(full, flush, 10)
    ref_out = [t.clone() for t in full()]
    # the form a deployment uses: views alternating over three streams (NMS + map-back of one view under the decode of the next), one graph
    replay, g_out = fusion.graph(pps, devh, n_streams=3)
    t_full = time_flushed(replay, flush, 10)
    replay()
    torch.cuda.synchronize()
    same = bool(all(torch.equal(a, b) for a, b in zip(g_out, ref_out)))
    t_wbf = time_flushed(lambda: fusion.fuse(), flush, 10)

    def tta_cpu():
        for b in range(2):
            bl, sl, ll = [], [], []
            for (hh, rr, ff, ss_) in views:
                d = oracle.yolo.non_max_suppression(oracle.yolo.decode_box([x[b:b + 1] for x in hh]), CONF, IOU)[0]
                bb, s_, lb = oracle.tta.map_back(d, rr, ff, float(ss_), 640.0, 640.0)
                bl.append(bb.numpy()); sl.append(s_.numpy()); ll.append(lb.numpy())
            oracle.wbf.weighted_boxes_fusion(bl, sl, ll, None, 0.55, 0.001)
    cpu_t = _best(tta_cpu, 1)
    out["cfg5"] = {"workload": "cfg5: YOLOv5s TTA 6 views (544/640/768 x {id,hflip}) batch 64: 6x(decode+filter+NMS) + map-back + Weighted Boxes Fusion",
                   "ms": t_full, "img_s": B / t_full * 1e3, "roofline": roof(nbytes, t_full, peak), "wbf_alone_ms": t_wbf,
                   "form": "TTAFusion.graph: views alternating over 3 CUDA streams, one CUDA graph", "eager_view_after_view_ms": t_eager,
                   "graph_equals_eager": same,
                   "fused_per_img": float(fusion.wbf.oc.float().mean()),
                   "cpu_baseline": {"value": 2 / cpu_t, "unit": "img/s", "cores": cpu_cores, "kind": "port", "sample": "2 images x 6 views, best of 1 after 1 warm-up"}}
    out["timing"] = "CUDA events around every iteration, a 256 MB buffer written between iterations (L2 flush, not timed)"
    return out


# ------------------------------------------------------------------------------------------------ main arm
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import torch
    import torch.distributed as dist
    from heltondetection_b200 import synth, yolo, _lib
    from heltondetection_b200 import dist as hd_dist

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = hd_dist.bind_to_gpu_numa(local)   # before any pinned allocation (also at N=1: the e2e leg reads pinned host memory)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    strong = args.scaling == "strong"
    if strong and args.batch % world:
        raise SystemExit(f"--batch {args.batch} must be divisible by the number of ranks ({world}) for equal shards")
    Bl = args.batch // world if strong else args.batch          # images this rank processes per step
    imgs_per_step = args.batch if strong else args.batch * world
    torch.set_num_threads(max(1, (len(os.sched_getaffinity(0)) or 8) // (1 if numa is not None else max(world, 1))))
    heads_cpu, _ = synth.yolo_heads(args.batch, IMG, NC, G, 1235 + (0 if strong else rank))
    if strong:      # pool slot j holds the shard of rank (rank + j) % world: slot 0 is this rank's own slice of the global batch
        shards = [hd_dist.shard_slice(args.batch, (rank + j) % world, world) for j in range(world)]
        pool_cpu = [[h[s] for h in heads_cpu] for s in shards]
    else:
        pool_cpu = [heads_cpu]
    pool = [[h.to(dev) for h in hs] for hs in pool_cpu]
    dense = args.mode == "dense"
    depth = args.depth if args.depth > 0 else (4 if Bl >= 128 else 8)

    # multi-GPU: the NMS kernels store every kept row into every rank's gather buffer (symmetric memory, posted NVLink stores) --
    # the all-gather is fused into the kernel epilogue; one cross-GPU barrier per step on the step's stream.
    # Fallback if symmetric memory is unavailable: pack + NCCL all_gather_into_tensor on a side stream.
    gather, peer, gather_mode, gather_parity = None, None, "none", None
    if world > 1:
        try:
            peer = hd_dist.PeerDetectionBuffers(Bl, MAX_DET, dev, slots=2 * depth)
            gather_mode = "fused: NMS kernels write into every peer's gather buffer over NVLink (symmetric memory) + per-step device barrier"
        except Exception as e:  # noqa: BLE001
            peer = None
            gather = hd_dist.DetectionGather(Bl, MAX_DET, dev)
            gather_mode = f"NCCL all_gather of padded detections on a side stream (symmetric memory unavailable: {type(e).__name__})"
    # the K timed steps are whole graph cycles: ceil(K/192) cycles of K // cycles steps (a remainder of < cycles steps runs step by step)
    n_cycles = max(1, -(-args.steps // 192))
    cycle_steps = max(1, args.steps // n_cycles)
    pipe = yolo.PostprocessPipeline(pool, depth=depth, peer=peer, device=dev, cycle_graph=True, min_cycle=cycle_steps, cycle_exact=True,
                                    conf_thres=CONF, iou_thres=IOU, max_det=MAX_DET, dense_read=dense)

    def run_step(k):
        det, cnt, _ = pipe.step(k)
        if gather is not None:
            with torch.cuda.stream(pipe.stream_of(k)):
                gather(det, cnt)

    def run_steps(n):   # whole cycles of steps are one graph launch each (fused-gather path); the rest step by step
        if gather is None:
            pipe.run(0, n)
        else:
            for k in range(n):
                run_step(k)

    def join():
        pipe.join()
        if gather is not None:
            gather.finish()

    if peer is not None:
        # one step through the fused gather, compared with pack + NCCL all_gather of the same detections
        pipe.fork(); pipe.step(0); pipe.join()
        torch.cuda.synchronize()
        det_l, cnt_l = peer.local(0)
        ref = hd_dist.DetectionGather(Bl, MAX_DET, dev)
        slot = ref(det_l.contiguous(), cnt_l.contiguous())
        ref_det, ref_cnt = ref.result(slot)
        torch.cuda.synchronize()
        got_det, got_cnt = peer.gathered(0)
        # rows at and beyond an image's count are unspecified in the peers' copies: compare the counts and the valid rows
        valid = (torch.arange(MAX_DET, device=dev)[None, :] < ref_cnt[:, None])[..., None]
        same = torch.equal(got_cnt, ref_cnt) and torch.equal(torch.where(valid, got_det, 0.0), torch.where(valid, ref_det.reshape(got_det.shape), 0.0))
        ok = torch.tensor([int(same)], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        gather_parity = bool(ok.item())

    sampler = ClockSampler(local)
    sampler.start()
    W = max(args.warmup, 3)
    pipe.fork()
    run_steps(W)
    join()
    if gather is None and pipe.cycle is not None:   # plus one untimed replay of the cycle graph (its first launch uploads the graph)
        pipe.fork()
        pipe.run(0, pipe.cycle_len)
        join()
    torch.cuda.synchronize()
    # ---------------- timed region: K pipelined steps, device-resident inputs (2.19 GB pool per rank >> 126 MB L2)
    K = args.steps
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.recording = True
    ev0.record()
    pipe.fork()
    run_steps(K)
    join()                                    # every step (and its gather) completes inside the timed region
    ev1.record()
    torch.cuda.synchronize()
    sampler.recording = False
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = imgs_per_step * K / (total_ms * 1e-3)

    # un-pipelined step time on one stream (explains the pipelined figure; not the headline)
    ks = min(K, 50)
    es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    es0.record()
    for k in range(ks):
        pipe.replays[k % pipe.n_graphs]()
    es1.record()
    torch.cuda.synchronize()
    serial_ms = es0.elapsed_time(es1) / ks

    # ---------------- e2e: public API with HOST (pinned) inputs; host<->device traffic inside the timed region, every step
    own_cpu = pool_cpu[0]
    pinned = [h.contiguous().pin_memory() for h in own_cpu]
    det_h = [torch.empty((Bl, MAX_DET, 6), dtype=torch.float32).pin_memory() for _ in range(2 * depth)]
    cnt_h = [torch.empty((Bl,), dtype=torch.int32).pin_memory() for _ in range(2 * depth)]
    stages = [[torch.empty_like(h, device=dev) for h in own_cpu] for _ in range(depth)]
    pipe_copy = yolo.PostprocessPipeline(stages, depth=depth, peer=peer, device=dev, conf_thres=CONF, iou_thres=IOU, max_det=MAX_DET, dense_read=False)
    e2e_steps = max(args.e2e_steps, 2) * (world if strong else 1)

    def e2e_loop(p, h2d):
        p.fork()
        for k in range(e2e_steps):
            s = p.stream_of(k)
            if h2d:
                with torch.cuda.stream(s):
                    for s_, p_ in zip(stages[k % depth], pinned):
                        s_.copy_(p_, non_blocking=True)
            det, cnt, _ = p.step(k)
            with torch.cuda.stream(s):
                det_h[k % (2 * depth)].copy_(det, non_blocking=True)
                cnt_h[k % (2 * depth)].copy_(cnt, non_blocking=True)
        p.join()
        torch.cuda.synchronize()

    def time_e2e(p, h2d):
        e2e_loop(p, h2d)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_loop(p, h2d)
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return imgs_per_step * e2e_steps / float(te.item())

    e2e_a = time_e2e(pipe_copy, True)
    det_a, cnt_a = det_h[0].clone(), cnt_h[0].clone()
    try:
        pipe_zc = yolo.PostprocessPipeline([pinned], depth=depth, peer=peer, device=dev, conf_thres=CONF, iou_thres=IOU, max_det=MAX_DET, dense_read=False)
        e2e_b = time_e2e(pipe_zc, False)
        zc_ok = bool(torch.equal(det_a, det_h[0]) and torch.equal(cnt_a, cnt_h[0]))
    except RuntimeError:
        e2e_b, zc_ok = 0.0, False
    full_bytes = sum(h.numel() * 4 for h in own_cpu) * world
    gate = math.log(CONF / (1 - CONF)) - 0.01
    zc_bytes = 0
    for h in own_cpu:                        # bytes the zero-copy kernel pulls: objectness planes + surviving 32-byte sectors
        Bn, Ctot, H, Wd = h.shape
        o = h.reshape(Bn, 3, Ctot // 3, H * Wd)[:, :, 4]
        pad = (-o.shape[-1]) % 8
        o = torch.nn.functional.pad(o, (0, pad), value=-100.0).view(Bn, 3, -1, 8)
        alive = (o > gate).any(-1)           # 32-byte sectors (8 cells) that hold a possible survivor
        zc_bytes += h.shape[0] * 3 * H * Wd * 4 + int(alive.sum()) * (Ctot // 3 - 1) * 32
    zc_bytes *= world                        # all ranks (shards are statistically alike)
    d2h = (det_h[0].numel() * 4 + cnt_h[0].numel() * 4) * world
    e2e_zero = {"value": e2e_b, "unit": "img/s", "h2d_bytes_per_step": zc_bytes, "d2h_bytes_per_step": d2h, "matches_full_copy": zc_ok,
                "path": "zero-copy: the decode kernel reads the pinned host tensors over PCIe (objectness planes + 32 B sectors of possible survivors)"}
    e2e_full = {"value": e2e_a, "unit": "img/s", "h2d_bytes_per_step": full_bytes, "d2h_bytes_per_step": d2h,
                "path": "cudaMemcpyAsync of the full head tensors from pinned memory, then the device path"}
    e2e = dict(e2e_zero if (e2e_b > e2e_a and zc_ok) else e2e_full)
    e2e["pipelined_steps"] = depth

    extra = {"e2e_zero_copy": e2e_zero, "e2e_full_copy": e2e_full, "step_ms_serial": serial_ms, "numa_node": numa}
    if dense:   # sparse (objectness-tile skip) variant, reported beside the dense headline
        pps = yolo.PostprocessPipeline(pool, depth=depth, peer=None, device=dev, conf_thres=CONF, iou_thres=IOU, max_det=MAX_DET, dense_read=False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pps.fork()
        for k in range(5):
            pps.step(k)
        pps.join()
        torch.cuda.synchronize()
        e0.record()
        pps.fork()
        for k in range(K):
            pps.step(k)
        pps.join()
        e1.record()
        torch.cuda.synchronize()
        ts = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        extra["sparse_skip_value"] = imgs_per_step * K / (float(ts.item()) * 1e-3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, copy)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        achieved = BYTES_PER_IMG * Bl * K / (total_ms * 1e-3) / 1e9       # per GPU, over the whole pipelined timed region
        traffic, traffic_src = None, "no ncu capture of this round found under profiles/"
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
            ent = tj.get(f"yolo_decode_filter_kernel:B={Bl}:{args.mode}")
            if ent:
                traffic, traffic_src = ent["dram_read_bytes"] + ent["dram_write_bytes"], ent["source"]
        except Exception:
            pass
        out = {
            "metric": "post-process images/sec", "value": value, "unit": "img/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": imgs_per_step, "batch_per_gpu": Bl, "conf_thres": CONF, "iou_thres": IOU,
                       "max_det": MAX_DET, "read_mode": args.mode,
                       "l2": f"every rank rotates through a pool of {len(pool)} distinct shard inputs = {len(pool) * Bl * BYTES_PER_IMG / 1e9:.2f} GB per rank, larger than the 126 MB L2",
                       "launch": (f"one hd_yolo_postprocess call per step, steps pipelined {depth} deep over CUDA streams; "
                                  + (f"cycles of {pipe.cycle_len} steps replayed as one CUDA graph" if pipe.cycle is not None else
                                     f"one CUDA-graph replay per step (cycle graph unavailable: {getattr(pipe, 'cycle_error', 'off')})")),
                       "parallelism": f"image-sharded x{world}" + ("" if strong else " (weak: own batch per rank)"), "gather": gather_mode},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": int(pipe.launches_per_step) * K,
            "gather_parity": gather_parity,
            "roofline": {"kernel": "yolo_decode_filter_kernel (decode+sigmoid+filter+compaction); measured over the whole pipelined step, NMS kernels included",
                         "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "frac_of_nominal_8000": achieved / 8000.0, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": BYTES_PER_IMG * Bl, "per": "GPU"},
        }
        out.update(extra)
        if world == 1:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            n = min(args.cpu_sample, args.batch)
            v = cpu_reference_rate(heads_cpu, n)
            out["cpu_baseline"] = {"value": v, "unit": "img/s", "cores": torch.get_num_threads(), "kind": "port",
                                   "sample": f"{n} images of the same batch, best of 3 after 1 warm-up"}
            if not args.no_configs:
                del pool, pipe, pipe_copy, stages
                torch.cuda.empty_cache()
                try:
                    out["configs"] = extra_configs(dev, peak, cores)
                except Exception as e:  # noqa: BLE001  (the headline line must still be printed)
                    out["configs"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

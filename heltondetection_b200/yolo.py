"""YOLOv5 post-process host API (drop-in names of the reference lineage, README.md:9).

``decode_box`` / ``non_max_suppression`` keep the lineage signatures;
``YoloPostprocessor`` is the fast path (raw heads -> detections, decode kernel + NMS kernels, no host sync).
"""
import ctypes as C
import os
import torch
from . import _lib

DEFAULT_ANCHORS = (
    ((10, 13), (16, 30), (33, 23)),
    ((30, 61), (62, 45), (59, 119)),
    ((116, 90), (156, 198), (373, 326)),
)
DEFAULT_STRIDES = (8, 16, 32)
_CLASS_MODES = {"agnostic": _lib.NMS_AGNOSTIC, "exact": _lib.NMS_CLASS_EXACT, "offset": _lib.NMS_CLASS_OFFSET}


def _is_pinned_host(x):
    return (not x.is_cuda) and x.is_pinned()


_HALF_FLAGS = {torch.float16: 4, torch.bfloat16: 8}   # HD_FLAG_IN_F16 / HD_FLAG_IN_BF16
_FLAG_IN_NHWC = 128


def _is_nhwc(x):
    """channels_last storage of a [B,C,H,W] head (and not at the same time plain contiguous)"""
    return x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous()


def _dtype_flag(keep):
    if keep[0].dtype == torch.float32 and all(_is_nhwc(x) for x in keep):
        return _FLAG_IN_NHWC
    return _HALF_FLAGS.get(keep[0].dtype, 0)


def _levels(outputs, anchors, strides, host_ok=False, half_ok=False):
    if len(outputs) != len(anchors) or len(outputs) != len(strides):
        raise RuntimeError("outputs, anchors and strides must have one entry per level")
    A = len(anchors[0])
    B, Ctot = outputs[0].shape[0], outputs[0].shape[1]
    if Ctot % A != 0 or Ctot // A < 6:
        raise RuntimeError(f"head channels {Ctot} are not A*(5+nc) with A={A}")
    nc = Ctot // A - 5
    arr = (_lib.YoloLevel * len(outputs))()
    keep = []
    total = 0
    for l, (x, anc, s) in enumerate(zip(outputs, anchors, strides)):
        if not (host_ok and _is_pinned_host(x)):
            _lib.require_cuda(x)
        if x.dim() != 4 or x.shape[0] != B or x.shape[1] != Ctot or len(anc) != A:
            raise RuntimeError(f"level {l}: expected [B={B}, {Ctot}, H, W], got {tuple(x.shape)}")
        if half_ok and x.dtype in _HALF_FLAGS and x.dtype == outputs[0].dtype:
            x = x if x.is_contiguous() else x.contiguous()   # 16-bit heads are widened to fp32 inside the kernel
        elif half_ok and x.dtype == torch.float32 and x.is_cuda and all(_is_nhwc(o) and o.dtype == torch.float32 for o in outputs):
            pass                                             # channels_last heads are read in place by the NHWC kernel
        else:
            x = _lib.f32c(x)
        keep.append(x)
        arr[l].data = x.data_ptr()
        arr[l].H, arr[l].W, arr[l].stride = x.shape[2], x.shape[3], float(s)
        for a, (w, h) in enumerate(anc):
            arr[l].anchor_wh[2 * a], arr[l].anchor_wh[2 * a + 1] = float(w), float(h)
        total += A * x.shape[2] * x.shape[3]
    return arr, keep, B, A, nc, total


@_lib.on_device
def decode_box(outputs, anchors=DEFAULT_ANCHORS, strides=DEFAULT_STRIDES):
    """list of [B, A*(5+nc), H, W] -> [B, sum(A*H*W), 5+nc] (cx,cy,w,h,obj,cls..) px (A.1)."""
    arr, keep, B, A, nc, total = _levels(outputs, anchors, strides)
    pred = torch.empty((B, total, 5 + nc), dtype=torch.float32, device=keep[0].device)
    _lib.check(_lib.lib().hd_yolo_decode(arr, len(keep), B, A, nc, _lib.ptr(pred), _lib.stream()))
    return pred


class _Buffers:
    """Candidate + NMS buffers for (B, cap, max_det) on one device; reused across calls."""

    def __init__(self, B, cap, max_det, device):
        self.B, self.cap, self.max_det = B, cap, max_det
        self.box = torch.empty((B, cap, 4), dtype=torch.float32, device=device)
        self.score = torch.empty((B, cap), dtype=torch.float32, device=device)
        self.cls = torch.empty((B, cap), dtype=torch.int32, device=device)
        self.anchor = torch.empty((B, cap), dtype=torch.int32, device=device)
        self.count = torch.zeros((B,), dtype=torch.int32, device=device)
        self.ws_bytes = _lib.lib().hd_sort_nms_workspace_size(B, cap)
        self.ws = torch.empty((self.ws_bytes,), dtype=torch.uint8, device=device)
        self.det = torch.zeros((B, max_det, 6), dtype=torch.float32, device=device)
        self.idx = torch.zeros((B, max_det), dtype=torch.int64, device=device)
        self.out_count = torch.zeros((B,), dtype=torch.int32, device=device)


def _run_nms(buf, iou_thres, class_mode, max_wh, max_nms):
    _lib.check(_lib.lib().hd_sort_nms_batched(
        _lib.ptr(buf.box), _lib.ptr(buf.score), _lib.ptr(buf.cls), _lib.ptr(buf.anchor), _lib.ptr(buf.count), 0,
        buf.B, buf.cap, float(iou_thres), class_mode, float(max_wh), int(max_nms), buf.max_det,
        _lib.ptr(buf.det), _lib.ptr(buf.idx), _lib.ptr(buf.out_count), _lib.ptr(buf.ws), buf.ws_bytes, _lib.stream()))


class YoloPostprocessor:
    """Raw-head post-process: decode+filter+compaction kernel, then per-image sort+NMS kernels.

    __call__(outputs) -> (det [B,max_det,6] = x1,y1,x2,y2,conf,cls ; count [B] int32 ; idx [B,max_det] anchor ids)
    all on the device, padded, with no host synchronisation (CUDA-graph capturable).  `outputs` may be fp32, or fp16 /
    bf16 heads (an autocast model): the kernel widens them to fp32 on load and computes in fp32, i.e. the result equals
    the fp32 path run on ``outputs.float()``, at half the HBM traffic.  fp32 heads in torch.channels_last memory (what a
    channels_last model's final 1x1 conv produces) are read in place by the NHWC kernel; same candidates as the NCHW path."""

    def __init__(self, anchors=DEFAULT_ANCHORS, strides=DEFAULT_STRIDES, conf_thres=0.25, iou_thres=0.45,
                 agnostic=False, max_det=300, max_nms=30000, max_wh=7680.0, class_mode="offset", ge=False,
                 dense_read=False, one_call=True, device=None, multi_label=False, multi_label_cap=4):
        """one_call: a single C-ABI call (hd_yolo_postprocess, internal workspace); False issues the decode and the
        NMS entry points separately (same kernels, candidate buffers visible).  device: where the outputs live;
        needed when `outputs` are pinned HOST tensors, which the kernel then reads directly over PCIe (zero-copy:
        only the sectors of possible survivors are fetched)."""
        # multi_label needs candidate buffers larger than one slot per anchor: up to multi_label_cap (anchor, class) pairs per anchor
        # on average (an image with more is truncated arbitrarily -- `overflowed()` tells); that uses the two-call path, whose
        # capacity the caller chooses, instead of the one-call entry point, which sizes its workspace for one slot per anchor
        self.multi_label, self.multi_cap = bool(multi_label), max(1, int(multi_label_cap))
        self.one_call, self.device = bool(one_call) and not multi_label, (torch.device(device) if device is not None else None)
        self._fkey = None
        self.anchors, self.strides = anchors, strides
        self.conf_thres, self.iou_thres = float(conf_thres), float(iou_thres)
        self.max_det, self.max_nms, self.max_wh = int(max_det), int(max_nms), float(max_wh)
        self.class_mode = _lib.NMS_AGNOSTIC if agnostic else _CLASS_MODES[class_mode]
        # multi_label (ultralytics eval setting): every (anchor, class) pair over the threshold is a candidate; idx = anchor*nc + class
        self.flags = ((_lib.FLAG_CONF_GE if ge else 0) | (_lib.FLAG_DENSE_READ if dense_read else 0)
                      | (_lib.FLAG_MULTI_LABEL if multi_label else 0))
        self._buf = None

    def buffers(self, B, cap, device):
        b = self._buf
        if b is None or b.B != B or b.cap != cap or b.box.device != device:
            b = self._buf = _Buffers(B, cap, self.max_det, device)
        return b

    def _out_device(self, keep):
        if self.device is not None:
            return self.device
        if keep[0].is_cuda:
            return keep[0].device
        return torch.device("cuda", torch.cuda.current_device())

    def _one_call(self, arr, keep, B, A, nc, total, peer=None, slot=0):
        dev = self._out_device(keep)
        key = (B, total, dev)
        if self._fkey != key:
            L = _lib.lib()
            self.f_ws_bytes = L.hd_yolo_postprocess_workspace_size(B, total)
            self.f_ws = torch.empty((self.f_ws_bytes,), dtype=torch.uint8, device=dev)
            self.f_det = torch.zeros((B, self.max_det, 6), dtype=torch.float32, device=dev)
            self.f_idx = torch.zeros((B, self.max_det), dtype=torch.int64, device=dev)
            self.f_count = torch.zeros((B,), dtype=torch.int32, device=dev)
            self._fkey = key
        det, count, rep = self.f_det, self.f_count, None
        if peer is not None:   # outputs go to this rank's slice of the gather buffer and, replicated, to every peer's
            det, count, rep = peer.local(slot)[0], peer.local(slot)[1], peer.replicas(slot)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().hd_yolo_postprocess_replicated(
                arr, len(keep), B, A, nc, self.conf_thres, self.iou_thres, self.flags | self._dflag, self.class_mode, self.max_wh, self.max_nms,
                self.max_det, _lib.ptr(det), _lib.ptr(self.f_idx), _lib.ptr(count), rep, _lib.ptr(self.f_ws), self.f_ws_bytes,
                _lib.stream()))
        return det, count, self.f_idx

    def __call__(self, outputs, peer=None, slot=0):
        """peer: a dist.PeerDetectionBuffers -- the NMS kernels then store every kept row into this rank's slice of the
        gather buffer of EVERY rank (posted NVLink stores), i.e. the all-gather is fused into the kernel epilogue."""
        # The level table only depends on the pointers/shapes, so it is reused while they do not change -- but ONLY when the
        # kernel reads every level in place.  A level that had to be converted (non-contiguous view, mixed layouts) points at
        # a private copy, which a later call with the same pointer and shape would read stale: those are converted every call.
        sig = tuple((x.data_ptr(), tuple(x.shape), x.dtype, tuple(x.stride())) for x in outputs)
        if getattr(self, "_sig", None) != sig:
            self._lv = _levels(outputs, self.anchors, self.strides, host_ok=True, half_ok=True)
            self._in_place = all(k is o for k, o in zip(self._lv[1], outputs))
            self._sig = sig if self._in_place else None
        arr, keep, B, A, nc, total = self._lv
        self._dflag = _dtype_flag(keep)
        if self.one_call:
            return self._one_call(arr, keep, B, A, nc, total, peer, slot)
        if peer is not None:
            raise RuntimeError("replicated (peer) outputs need the one-call path")
        dev = self._out_device(keep)
        buf = self.buffers(B, min(total * min(nc, self.multi_cap), 500000) if self.multi_label else total, dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().hd_yolo_decode_filter(
                arr, len(keep), B, A, nc, self.conf_thres, self.flags | self._dflag, _lib.ptr(buf.box), _lib.ptr(buf.score),
                _lib.ptr(buf.cls), _lib.ptr(buf.anchor), _lib.ptr(buf.count), buf.cap, _lib.stream()))
            _run_nms(buf, self.iou_thres, self.class_mode, self.max_wh, self.max_nms)
        return buf.det, buf.out_count, buf.idx

    def overflowed(self):
        """multi_label: True if some image of the last call had more candidates than the buffer holds (one host sync)"""
        b = self._buf
        return bool(b is not None and int(b.count.max()) > b.cap)

    def graph(self, outputs, warmup=3, peer=None, slot=0):
        """Capture one post-process of `outputs` (fixed buffers) into a CUDA graph; returns (replay, det, count, idx).
        replay() re-runs the captured kernels on whatever the input buffers hold -- no per-call host work."""
        for _ in range(warmup):
            self(outputs, peer, slot)
        if not self._in_place:
            raise RuntimeError("graph() needs heads the kernel reads in place (contiguous NCHW, channels_last fp32 or contiguous 16-bit): "
                               "a converted copy would be frozen into the graph")
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            det, count, idx = self(outputs, peer, slot)
        return g.replay, det, count, idx

    @_lib.on_device
    def candidates(self, outputs):
        """decode+filter only -> per image (cand [n,6], anchor idx [n]) sorted by anchor index (test helper)."""
        arr, keep, B, A, nc, total = _levels(outputs, self.anchors, self.strides, half_ok=True)
        buf = self.buffers(B, total, keep[0].device)
        _lib.check(_lib.lib().hd_yolo_decode_filter(
            arr, len(keep), B, A, nc, self.conf_thres, self.flags | _dtype_flag(keep), _lib.ptr(buf.box), _lib.ptr(buf.score),
            _lib.ptr(buf.cls), _lib.ptr(buf.anchor), _lib.ptr(buf.count), buf.cap, _lib.stream()))
        return _collect_candidates(buf)


def _collect_candidates(buf):
    out = []
    for b, n in enumerate(buf.count.tolist()):
        n = min(n, buf.cap)
        o = torch.argsort(buf.anchor[b, :n])
        c = torch.cat((buf.box[b, :n], buf.score[b, :n, None], buf.cls[b, :n, None].float()), 1)[o]
        out.append((c, buf.anchor[b, :n][o].long()))
    return out


def _slice(det, count, idx=None, group_by_class=False):
    counts = count.tolist()  # the one host sync of the list-returning drop-in
    dets = [det[b, :n] for b, n in enumerate(counts)]
    idxs = None if idx is None else [idx[b, :n] for b, n in enumerate(counts)]
    if group_by_class:   # lineage (bubbliiiing) order: class by class, score-descending inside a class
        for b, d in enumerate(dets):
            o = torch.sort(d[:, 5], stable=True)[1]
            dets[b] = d[o]
            if idxs is not None:
                idxs[b] = idxs[b][o]
    return dets if idx is None else (dets, idxs)


def postprocess(outputs, conf_thres=0.25, iou_thres=0.45, return_index=False, group_by_class=False, **kw):
    """Raw heads -> list of [k,6] detections (fast path).  group_by_class=True with class_mode="exact", ge=True and a large
    max_det reproduces the bubbliiiing-lineage output (per-class nms, results concatenated class by class)."""
    det, count, idx = YoloPostprocessor(conf_thres=conf_thres, iou_thres=iou_thres, **kw)(outputs)
    return _slice(det, count, idx, group_by_class) if return_index else _slice(det, count, None, group_by_class)


@_lib.on_device
def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, agnostic=False, max_det=300, max_nms=30000,
                        max_wh=7680.0, class_mode="offset", ge=False, return_index=False):
    """Lineage signature: decoded prediction [B, N, 5+nc] -> list of [k,6] (xyxy, conf, cls) (A.2)."""
    _lib.require_cuda(prediction)
    if prediction.dim() != 3 or prediction.shape[2] < 6:
        raise RuntimeError(f"prediction should be [B, N, 5+nc], got {tuple(prediction.shape)}")
    pred = _lib.f32c(prediction)
    B, N, no = pred.shape
    buf = _Buffers(B, max(N, 1), max_det, pred.device)
    flags = _lib.FLAG_CONF_GE if ge else 0
    _lib.check(_lib.lib().hd_yolo_filter_pred(
        _lib.ptr(pred), B, N, no - 5, float(conf_thres), flags, _lib.ptr(buf.box), _lib.ptr(buf.score),
        _lib.ptr(buf.cls), _lib.ptr(buf.anchor), _lib.ptr(buf.count), buf.cap, _lib.stream()))
    mode = _lib.NMS_AGNOSTIC if agnostic else _CLASS_MODES[class_mode]
    _run_nms(buf, iou_thres, mode, max_wh, max_nms)
    return _slice(buf.det, buf.out_count, buf.idx) if return_index else _slice(buf.det, buf.out_count)


class PostprocessPipeline:
    """Throughput mode: a `depth`-deep software pipeline of YoloPostprocessor steps over CUDA streams.

    Step k runs on stream k % depth with its own workspace and output buffers, so the (latency-bound, few-SM) NMS kernels of
    step k overlap the (HBM-bound) decode kernel of step k+1, and the drain of one decode grid is filled by the head of the next.
    Every step is one CUDA-graph replay of the single C-ABI call.  `pool` is a list of input sets (each the list of head tensors of
    one batch); step k reads pool[k % len(pool)].  With `peer` (dist.PeerDetectionBuffers, slots a multiple of depth; 2*depth keeps
    the barrier off the critical path) the NMS kernels store their rows into every rank's gather buffer; the cross-GPU barrier that
    marks a step complete runs on a side stream, and a gather-buffer slot is only rewritten after the barrier of its previous use.

        pipe = PostprocessPipeline(pool, depth=2, conf_thres=0.25)
        pipe.fork(); [pipe.step(k) for k in range(K)]; pipe.join()      # outputs of step k: pipe.outputs(k)
    """

    def __init__(self, pool, depth=2, peer=None, device=None, cycle_graph=False, min_cycle=48, cycle_exact=False, **pp_kwargs):
        import math
        self.pool, self.depth, self.peer, self.min_cycle, self.cycle_exact = list(pool), int(depth), peer, min_cycle, bool(cycle_exact)
        first = self.pool[0][0]
        self.device = torch.device(device) if device is not None else first.device
        if peer is not None and (peer.slots < self.depth or peer.slots % self.depth):
            raise RuntimeError("peer buffers need a multiple of `depth` slots")
        self.streams = [torch.cuda.Stream(self.device) for _ in range(self.depth)]
        self.pps = [YoloPostprocessor(device=self.device, **pp_kwargs) for _ in range(self.depth)]
        self.n_slots = peer.slots if peer is not None else self.depth
        self.n_graphs = len(self.pool) * self.n_slots // math.gcd(len(self.pool), self.n_slots)
        if peer is not None:
            self.side = torch.cuda.Stream(self.device)
            self._stepped = [torch.cuda.Event() for _ in range(self.n_slots)]
            self._gathered = [torch.cuda.Event() for _ in range(self.n_slots)]
            self._used = [False] * self.n_slots
        self.replays, self.outs = [], []
        L = _lib.lib()
        with torch.cuda.device(self.device):
            for g in range(self.n_graphs):
                c0 = L.hd_debug_launch_count()
                rp, det, cnt, idx = self.pps[g % self.depth].graph(self.pool[g % len(self.pool)], warmup=2, peer=peer, slot=g % self.n_slots)
                # launches of one step = what one capture recorded (2 warm-up calls + 1 captured call were counted)
                self.launches_per_step = (L.hd_debug_launch_count() - c0) // 3
                self.replays.append(rp)
                self.outs.append((det, cnt, idx))
        self._ev = torch.cuda.Event()
        self._dirty, self._side_dirty = [True] * self.depth, True              # streams that may hold work a cycle replay has not waited for
        self._fork_pending, self._cycle_pending, self._side_cycle_pending = [False] * self.depth, [False] * self.depth, False
        self.cycle = None
        if cycle_graph:
            self._capture_cycle()

    def _capture_cycle(self):
        """One CUDA graph holding a whole cycle of n_graphs consecutive steps -- their streams, the gather-slot dependencies and
        the completion barriers become graph branches -- so that the host issues one launch per cycle instead of ~6 calls per
        step (at 32 images a step is ~40 us of GPU time: a Python loop issuing it step by step is host bound)."""
        dev, depth = self.device, self.depth
        # a cycle ends with the pipeline drained, so it should hold many steps: a multiple of the (input, workspace, gather slot)
        # period that is at least ~48 steps long -- or (cycle_exact) exactly min_cycle steps, for a caller that runs a fixed number
        # of steps and wants them to be whole cycles (a cycle always starts from step 0's input / workspace / slot assignment and
        # is fenced from its neighbours by full joins, so its length need not be a multiple of the period)
        if self.cycle_exact:
            self.cycle_len = max(1, int(self.min_cycle))
        else:
            self.cycle_len = self.n_graphs * max(1, -(-int(self.min_cycle) // self.n_graphs))
        main = torch.cuda.Stream(dev)
        streams = [torch.cuda.Stream(dev) for _ in range(depth)]
        side = torch.cuda.Stream(dev)
        g = torch.cuda.CUDAGraph()
        try:
            torch.cuda.synchronize(dev)
            with torch.cuda.device(dev), torch.cuda.graph(g, stream=main):
                start = torch.cuda.Event()
                start.record(main)
                for s in streams + [side]:
                    s.wait_event(start)
                gathered = {}
                for k in range(self.cycle_len):
                    s = streams[k % depth]
                    slot = k % self.n_slots
                    with torch.cuda.stream(s):
                        if slot in gathered:
                            s.wait_event(gathered[slot])
                        self.pps[k % depth](self.pool[k % len(self.pool)], self.peer, slot)
                        if self.peer is not None:
                            stepped = torch.cuda.Event()
                            stepped.record(s)
                            with torch.cuda.stream(side):
                                side.wait_event(stepped)
                                if not os.environ.get("HD_PIPE_NO_BARRIER"):     # developer ablation switch
                                    self.peer.barrier(channel=slot)
                                done = torch.cuda.Event()
                                done.record(side)
                            gathered[slot] = done
                for s in streams + [side]:
                    main.wait_stream(s)
            self.cycle, self._main = g, main
        except Exception as e:  # noqa: BLE001  (e.g. a barrier that cannot be captured): the per-step path remains
            self.cycle, self.cycle_error = None, f"{type(e).__name__}: {e}"[:200]
            torch.cuda.synchronize(dev)

    def run(self, k0, n_steps):
        """steps k0 .. k0+n_steps-1 with as few host calls as possible (whole cycles as one graph launch, the remainder step by
        step); the cycle graph is used when k0 is a multiple of the (input, workspace, slot) period.  Cross-stream dependencies
        are issued lazily: a cycle replay waits only for pipeline streams that received work since the last replay, and the
        streams wait for the replay when they are next used -- a run made of whole cycles costs one wait + one graph launch per
        cycle on the host (at 32-image shards the ~20 stream waits of the eager form delayed the launch by more than a step)."""
        k = k0
        if self.cycle is not None and k % self.n_graphs == 0:
            cur = torch.cuda.current_stream(self.device)
            while k0 + n_steps - k >= self.cycle_len:
                self._main.wait_stream(cur)
                for i, st in enumerate(self.streams):
                    if self._dirty[i]:
                        self._main.wait_stream(st)
                        self._dirty[i] = False
                if self.peer is not None and self._side_dirty:
                    self._main.wait_stream(self.side)
                    self._side_dirty = False
                with torch.cuda.stream(self._main):
                    self.cycle.replay()
                self._cycle_pending = [True] * self.depth     # later single steps follow the cycle (wait issued at their first use)
                if self.peer is not None:
                    self._side_cycle_pending = True
                    self._used = [True] * self.n_slots
                k += self.cycle_len
        while k < k0 + n_steps:
            self.step(k)
            k += 1

    def fork(self):
        """the pipeline streams wait for the work queued so far on the current stream (the wait itself is issued when a stream is
        next used; a cycle replay orders itself after the current stream anyway)"""
        self._ev.record()
        self._fork_pending = [True] * self.depth

    def _use_stream(self, i):
        st = self.streams[i]
        if self._fork_pending[i]:
            st.wait_event(self._ev)
            self._fork_pending[i] = False
        if self._cycle_pending[i]:
            st.wait_stream(self._main)
            self._cycle_pending[i] = False
        self._dirty[i] = True
        return st

    def step(self, k):
        g = k % self.n_graphs
        s = self._use_stream(g % self.depth)
        with torch.cuda.stream(s):
            if self.peer is not None:
                slot = g % self.n_slots
                if self._used[slot]:
                    s.wait_event(self._gathered[slot])      # the slot's previous contents have been gathered everywhere
                self.replays[g]()
                self._stepped[slot].record(s)
                if self._side_cycle_pending:
                    self.side.wait_stream(self._main)
                    self._side_cycle_pending = False
                self._side_dirty = True
                with torch.cuda.stream(self.side):           # completion barrier off the compute streams
                    self.side.wait_event(self._stepped[slot])
                    self.peer.barrier(channel=slot)
                    self._gathered[slot].record(self.side)
                self._used[slot] = True
            else:
                self.replays[g]()
        return self.outs[g]

    def outputs(self, k):
        return self.outs[k % self.n_graphs]

    def stream_of(self, k):
        """the stream step k runs on, ordered after the last fork() / cycle replay -- for the caller's own copies around the step"""
        return self._use_stream((k % self.n_graphs) % self.depth)

    def join(self):
        """the current stream waits for every pipeline stream"""
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            cur.wait_stream(s)
        if self.peer is not None:
            cur.wait_stream(self.side)
        if self.cycle is not None:
            cur.wait_stream(self._main)

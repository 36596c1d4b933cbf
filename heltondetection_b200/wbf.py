"""Weighted Boxes Fusion + TTA host API (README.md:19).

``weighted_boxes_fusion`` keeps the ensemble-boxes signature for one image; ``WbfBatched`` / ``TTAFusion``
are the batched, sync-free fast path used after the per-view YOLO post-process.
"""
import ctypes as C
import numpy as np
import torch
from . import _lib

_CONF = {"avg": _lib.WBF_AVG, "max": _lib.WBF_MAX, "box_and_model_avg": _lib.WBF_BOX_AND_MODEL_AVG,
         "absent_model_aware_avg": _lib.WBF_ABSENT_MODEL_AWARE_AVG}


class WbfBatched:
    """boxes [B,V,M,4], scores [B,V,M], labels [B,V,M] (float), counts [B,V] int32 ->
    (boxes [B,V*M,4] f32, scores [B,V*M] f64, labels [B,V*M] f32, count [B] int32).  The output buffers are reused by every call;
    rows of image b at and beyond count[b] are unspecified (whatever an earlier call left there)."""

    def __init__(self, num_labels, weights=None, iou_thr=0.55, skip_box_thr=0.0, conf_type="avg", allows_overflow=False,
                 rescale="len_weights"):
        """rescale: the 'avg' confidence rescale of ensemble-boxes >= 1.0.5, min(n, len(weights)) ("len_weights", default), or of
        older releases, min(n, sum(weights)) ("sum_weights"); identical for unit weights."""
        if conf_type not in _CONF:
            raise RuntimeError(f'Unknown conf_type: {conf_type}. Must be "avg", "max", "box_and_model_avg" or "absent_model_aware_avg"')
        if rescale not in ("len_weights", "sum_weights"):
            raise RuntimeError('rescale must be "len_weights" or "sum_weights"')
        self.num_labels, self.weights = int(num_labels), weights
        self.iou_thr, self.skip, self.overflow = float(iou_thr), float(skip_box_thr), int(allows_overflow)
        self.conf = _CONF[conf_type] | (_lib.WBF_RESCALE_SUM_WEIGHTS if rescale == "sum_weights" else 0)
        self._key = None

    def _alloc(self, B, V, M, dev):
        key = (B, V, M, dev)
        if self._key != key:
            L = _lib.lib()
            self.ws_bytes = L.hd_wbf_workspace_size(B, V, M, self.num_labels)
            self.ws = torch.empty((self.ws_bytes,), dtype=torch.uint8, device=dev)
            self.ob = torch.zeros((B, V * M, 4), dtype=torch.float32, device=dev)
            self.os = torch.zeros((B, V * M), dtype=torch.float64, device=dev)
            self.ol = torch.zeros((B, V * M), dtype=torch.float32, device=dev)
            self.oc = torch.zeros((B,), dtype=torch.int32, device=dev)
            self._key = key

    @_lib.on_device
    def __call__(self, boxes, scores, labels, counts):
        _lib.require_cuda(boxes, scores, labels, counts)
        B, V, M = scores.shape
        self._alloc(B, V, M, boxes.device)
        w = None
        if self.weights is not None:
            if len(self.weights) != V:
                raise RuntimeError(f"Incorrect number of weights {len(self.weights)}. Must be: {V}")
            w = (C.c_double * V)(*[float(x) for x in self.weights])
        # converted copies are bound to locals so that they outlive the launch (a temporary's block could be handed to the
        # next conversion by the caching allocator before the kernel has read it)
        bx, sc, lb, cn = _lib.f32c(boxes), _lib.f32c(scores), _lib.f32c(labels), counts.to(torch.int32).contiguous()
        _lib.check(_lib.lib().hd_wbf(_lib.ptr(bx), _lib.ptr(sc), _lib.ptr(lb),
                                     _lib.ptr(cn), B, V, M, self.num_labels, w, self.iou_thr, self.skip,
                                     self.conf, self.overflow, _lib.ptr(self.ob), _lib.ptr(self.os), _lib.ptr(self.ol), _lib.ptr(self.oc),
                                     _lib.ptr(self.ws), self.ws_bytes, _lib.stream()))
        return self.ob, self.os, self.ol, self.oc


def weighted_boxes_fusion(boxes_list, scores_list, labels_list, weights=None, iou_thr=0.55, skip_box_thr=0.0,
                          conf_type="avg", allows_overflow=False, device="cuda", rescale="len_weights"):
    """ensemble-boxes signature, one image: lists (one entry per model/view) of boxes [n,4] in [0,1], scores, labels
    -> (boxes [m,4], scores [m], labels [m]) numpy float64, sorted by fused score."""
    V = len(boxes_list)
    M = max([len(b) for b in boxes_list] + [1])
    bx = torch.zeros((1, V, M, 4), dtype=torch.float32)
    sc = torch.zeros((1, V, M), dtype=torch.float32)
    lb = torch.zeros((1, V, M), dtype=torch.float32)
    cnt = torch.zeros((1, V), dtype=torch.int32)
    mx = 0
    for v in range(V):
        n = len(boxes_list[v])
        cnt[0, v] = n
        if n:
            bx[0, v, :n] = torch.as_tensor(np.asarray(boxes_list[v], np.float32)).reshape(n, 4)
            sc[0, v, :n] = torch.as_tensor(np.asarray(scores_list[v], np.float32))
            lab = torch.as_tensor(np.asarray(labels_list[v], np.float32))
            lb[0, v, :n] = lab
            mx = max(mx, int(lab.max().item()))
    f = WbfBatched(mx + 1, weights, iou_thr, skip_box_thr, conf_type, allows_overflow, rescale)
    ob, os_, ol, oc = f(bx.to(device), sc.to(device), lb.to(device), cnt.to(device))
    m = int(oc.item())
    return ob[0, :m].double().cpu().numpy(), os_[0, :m].cpu().numpy(), ol[0, :m].double().cpu().numpy()


class TTAFusion:
    """Per-view detections -> fused detections.  views: list of (scale, hflip, view_w)."""

    def __init__(self, views, img_size, num_labels, max_det=300, **wbf_kw):
        self.views, self.img_h, self.img_w, self.M = views, float(img_size[0]), float(img_size[1]), int(max_det)
        self.wbf = WbfBatched(num_labels, **wbf_kw)
        self._key = None

    def _alloc(self, B, dev):
        if self._key != (B, dev):
            V, M = len(self.views), self.M
            self.bx = torch.zeros((B, V, M, 4), dtype=torch.float32, device=dev)
            self.sc = torch.zeros((B, V, M), dtype=torch.float32, device=dev)
            self.lb = torch.zeros((B, V, M), dtype=torch.float32, device=dev)
            self.cnt = torch.zeros((B, V), dtype=torch.int32, device=dev)
            self._key = (B, dev)

    @_lib.on_device
    def map_back(self, v, det, count):
        """det [B,max_det,6] / count [B] of view v (device) -> written into the WBF input slot v."""
        B = det.shape[0]
        self._alloc(B, det.device)
        scale, hflip, view_w = self.views[v]
        det, count = _lib.f32c(det), count.to(torch.int32).contiguous()   # locals: alive until the launch is queued
        _lib.check(_lib.lib().hd_tta_map_back(_lib.ptr(det), _lib.ptr(count), B, det.shape[1], float(scale), int(bool(hflip)),
                                              float(view_w), self.img_w, self.img_h, _lib.ptr(self.bx), _lib.ptr(self.sc), _lib.ptr(self.lb),
                                              _lib.ptr(self.cnt), len(self.views), v, self.M, _lib.stream()))

    def fuse(self):
        return self.wbf(self.bx, self.sc, self.lb, self.cnt)

    def run(self, postprocessors, view_outputs, n_streams=2):
        """Whole TTA step: view v's decode + NMS + map-back run on side stream v % n_streams, so the (small, latency-bound) NMS and
        map-back kernels of one view execute under the (HBM-bound) decode of the next; WBF follows the join on the current stream.
        postprocessors: one yolo.YoloPostprocessor per view (each owns its buffers); view_outputs: per view the list of head
        tensors.  Same results as calling the views one after the other; no host synchronisation, CUDA-graph capturable."""
        first = view_outputs[0][0]
        dev = first.device if first.is_cuda else postprocessors[0].device
        with torch.cuda.device(dev):
            self._alloc(int(first.shape[0]), dev)
            cur = torch.cuda.current_stream(dev)
            if getattr(self, "_streams_dev", None) != (dev, n_streams):
                self._side = [torch.cuda.Stream(device=dev) for _ in range(max(1, int(n_streams)))]
                self._streams_dev = (dev, n_streams)
            fork = torch.cuda.Event()
            fork.record(cur)
            for s in self._side:
                s.wait_event(fork)
            for v in range(len(self.views)):
                with torch.cuda.stream(self._side[v % len(self._side)]):
                    det, count, _ = postprocessors[v](view_outputs[v])
                    self.map_back(v, det, count)
            for s in self._side:
                ev = torch.cuda.Event()
                ev.record(s)
                cur.wait_event(ev)
            return self.fuse()

    def graph(self, postprocessors, view_outputs, n_streams=2, warmup=3):
        """Capture run() on fixed input buffers into one CUDA graph -> (replay, fused outputs)."""
        for _ in range(warmup):
            out = self.run(postprocessors, view_outputs, n_streams)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = self.run(postprocessors, view_outputs, n_streams)
        return g.replay, out

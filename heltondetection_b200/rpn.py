"""RPN proposal creation host API (README.md:8,63-65).

``RpnProposals`` is the fused fast path from the raw NCHW RPN heads (anchors generated in-kernel);
``ProposalCreator`` keeps the lineage call signature ``(loc, score, anchor, img_size, scale)`` on
already flattened per-image tensors.
"""
import math
import numpy as np
import torch
from . import _lib


def generate_anchor_base(base_size, ratios=(0.5, 1.0, 2.0), scales=(8.0,)):
    """[len(ratios)*len(scales), 4] fp32 (x1,y1,x2,y2) centred on 0: h=base*scale*sqrt(r), w=base*scale/sqrt(r)."""
    out = np.zeros((len(ratios) * len(scales), 4), np.float32)
    for i, r in enumerate(ratios):
        for j, sc in enumerate(scales):
            h = base_size * sc * math.sqrt(r)
            w = base_size * sc * math.sqrt(1.0 / r)
            out[i * len(scales) + j] = (-w / 2.0, -h / 2.0, w / 2.0, h / 2.0)
    return out


def set_mode(mode):
    """stage-2 kernel choice: 0 auto, 1 one CTA per image, 2 cluster of 8 CTAs per image; returns the previous mode."""
    return _lib.lib().hd_rpn_set_mode(int(mode))


def _levels(objectness, deltas, anchor_bases, strides, softmax):
    n = len(deltas)
    if not (len(objectness) == n == len(anchor_bases) == len(strides)):
        raise RuntimeError("objectness, deltas, anchor_bases and strides must have one entry per level")
    B, A = deltas[0].shape[0], deltas[0].shape[1] // 4
    arr = (_lib.RpnLevel * n)()
    keep = []
    for l in range(n):
        o, d = objectness[l], deltas[l]
        _lib.require_cuda(o, d)
        if d.dim() != 4 or d.shape[1] != 4 * A or o.shape[1] != (2 * A if softmax else A) or o.shape[2:] != d.shape[2:]:
            raise RuntimeError(f"level {l}: expected objectness [B,{2 * A if softmax else A},H,W] and deltas [B,{4 * A},H,W], "
                               f"got {tuple(o.shape)} and {tuple(d.shape)}")
        o, d = _lib.f32c(o), _lib.f32c(d)
        keep += [o, d]
        arr[l].objectness, arr[l].deltas = o.data_ptr(), d.data_ptr()
        arr[l].H, arr[l].W, arr[l].stride = d.shape[2], d.shape[3], float(strides[l])
        ab = np.asarray(anchor_bases[l], np.float32).reshape(-1)
        if ab.size != 4 * A:
            raise RuntimeError(f"level {l}: anchor_base must hold {A} boxes")
        for q in range(4 * A):
            arr[l].anchor_base[q] = float(ab[q])
    return arr, keep, B, A


class RpnProposals:
    """raw RPN heads -> (rois [B*n_post,5], count [B], scores [B,n_post], idx [B,n_post]); no host sync."""

    def __init__(self, anchor_bases, strides, img_size, nms_iou=0.7, n_pre_nms=12000, n_post_nms=2000, min_size=16.0,
                 score_mode="sigmoid", clamp_dwh=None, exact_math=False):
        """exact_math: sigmoid / softmax / exp in fp64 rounded once to fp32 (HD_RPN_EXACT_MATH) -- the proposals are then
        bit-reproducible against a CPU doing the same (oracle.rpn exact_math=True); default fp32 expf like the reference."""
        self.anchor_bases, self.strides, self.img_size = anchor_bases, strides, img_size
        self.nms_iou, self.n_pre, self.n_post, self.min_size = float(nms_iou), int(n_pre_nms), int(n_post_nms), float(min_size)
        self.flags = ((_lib.RPN_SOFTMAX if score_mode == "softmax" else 0) | (_lib.RPN_CLAMP_DWH if clamp_dwh is not None else 0)
                      | (_lib.RPN_EXACT_MATH if exact_math else 0))
        self.clamp = float(clamp_dwh) if clamp_dwh is not None else 0.0
        self._key = None

    def _alloc(self, B, N, dev):
        key = (B, N, dev)
        if self._key != key:
            L = _lib.lib()
            self.ws_bytes = L.hd_rpn_proposals_workspace_size(B, N, self.n_pre)
            self.ws = torch.empty((self.ws_bytes,), dtype=torch.uint8, device=dev)
            self.rois = torch.empty((B, self.n_post, 5), dtype=torch.float32, device=dev)
            self.scores = torch.empty((B, self.n_post), dtype=torch.float32, device=dev)
            self.idx = torch.empty((B, self.n_post), dtype=torch.int64, device=dev)
            self.count = torch.zeros((B,), dtype=torch.int32, device=dev)
            self._key = key

    @_lib.on_device
    def __call__(self, objectness, deltas):
        arr, keep, B, A = _levels(objectness, deltas, self.anchor_bases, self.strides, bool(self.flags & _lib.RPN_SOFTMAX))
        L = _lib.lib()
        N = L.hd_rpn_num_anchors(arr, len(deltas), A)
        self._alloc(B, N, keep[0].device)
        _lib.check(L.hd_rpn_proposals(arr, len(deltas), B, A, self.flags, float(self.img_size[0]), float(self.img_size[1]),
                                      self.min_size, self.clamp, self.n_pre, self.n_post, self.nms_iou, _lib.ptr(self.rois),
                                      _lib.ptr(self.scores), _lib.ptr(self.idx), _lib.ptr(self.count), _lib.ptr(self.ws),
                                      self.ws_bytes, _lib.stream()))
        return self.rois.view(B * self.n_post, 5), self.count, self.scores, self.idx

    @_lib.on_device
    def decode(self, objectness, deltas):
        """stage 1 only -> (boxes [B,N,4], scores [B,N], keys [B,N] int32 view of the sortable bits)."""
        arr, keep, B, A = _levels(objectness, deltas, self.anchor_bases, self.strides, bool(self.flags & _lib.RPN_SOFTMAX))
        L = _lib.lib()
        N = L.hd_rpn_num_anchors(arr, len(deltas), A)
        dev = keep[0].device
        boxes = torch.empty((B, N, 4), dtype=torch.float32, device=dev)
        scores = torch.empty((B, N), dtype=torch.float32, device=dev)
        keys = torch.empty((B, N), dtype=torch.int32, device=dev)
        _lib.check(L.hd_rpn_decode(arr, len(deltas), B, A, self.flags, float(self.img_size[0]), float(self.img_size[1]),
                                   self.min_size, self.clamp, _lib.ptr(boxes), _lib.ptr(scores), _lib.ptr(keys), _lib.stream()))
        return boxes, scores, keys


@_lib.on_device
def select_nms(boxes, scores, valid, n_pre, n_post, nms_iou):
    """stage 2 on explicit arrays: boxes [B,N,4], scores [B,N], valid [B,N] bool -> (rois [B,n_post,5], scores, idx, count)."""
    _lib.require_cuda(boxes, scores, valid)
    boxes, scores = _lib.f32c(boxes), _lib.f32c(scores)
    B, N = scores.shape
    # sortable keys exactly as the decode kernel builds them (-0 -> +0, NaN first, 0 = invalid)
    s = scores + 0.0
    bits = s.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    k = torch.where(bits >= 0x80000000, (~bits) & 0xFFFFFFFF, bits | 0x80000000)
    k = torch.where(torch.isnan(s), torch.full_like(k, 0xFFFFFFFF), k)
    k = torch.where(k == 0, torch.ones_like(k), k)
    k = torch.where(valid, k, torch.zeros_like(k))
    keys = (k & 0xFFFFFFFF).to(torch.int64)
    keys = torch.where(keys >= 2 ** 31, keys - 2 ** 32, keys).to(torch.int32).contiguous()
    L = _lib.lib()
    ws_bytes = L.hd_rpn_select_nms_workspace_size(B, N, n_pre)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=boxes.device)
    rois = torch.empty((B, n_post, 5), dtype=torch.float32, device=boxes.device)
    osc = torch.empty((B, n_post), dtype=torch.float32, device=boxes.device)
    idx = torch.empty((B, n_post), dtype=torch.int64, device=boxes.device)
    cnt = torch.zeros((B,), dtype=torch.int32, device=boxes.device)
    _lib.check(L.hd_rpn_select_nms(_lib.ptr(boxes), _lib.ptr(scores), _lib.ptr(keys), B, N, int(n_pre), int(n_post), float(nms_iou),
                                   _lib.ptr(rois), _lib.ptr(osc), _lib.ptr(idx), _lib.ptr(cnt), _lib.ptr(ws), ws_bytes, _lib.stream()))
    return rois, osc, idx, cnt


class ProposalCreator:
    """Lineage signature: __call__(loc [N,4], score [N], anchor [N,4], img_size (h,w), scale=1.) -> roi [k,4].

    The decode of explicit (already flattened) anchors is a handful of elementwise torch ops on the GPU; the
    top-k + sort + NMS runs in the hd_b200 kernel.  Short results are returned unpadded (A.3)."""

    def __init__(self, mode="test", nms_iou=0.7, n_train_pre_nms=12000, n_train_post_nms=2000, n_test_pre_nms=12000,
                 n_test_post_nms=2000, min_size=16):
        self.mode, self.nms_iou, self.min_size = mode, nms_iou, min_size
        self.n_train_pre_nms, self.n_train_post_nms = n_train_pre_nms, n_train_post_nms
        self.n_test_pre_nms, self.n_test_post_nms = n_test_pre_nms, n_test_post_nms

    @_lib.on_device
    def __call__(self, loc, score, anchor, img_size, scale=1.0, return_index=False):
        _lib.require_cuda(loc, score)
        n_pre, n_post = ((self.n_train_pre_nms, self.n_train_post_nms) if self.mode == "training"
                         else (self.n_test_pre_nms, self.n_test_post_nms))
        anchor = torch.as_tensor(anchor, dtype=torch.float32, device=loc.device)
        wa, ha = anchor[:, 2] - anchor[:, 0], anchor[:, 3] - anchor[:, 1]
        cxa, cya = anchor[:, 0] + 0.5 * wa, anchor[:, 1] + 0.5 * ha
        cx, cy = loc[:, 0] * wa + cxa, loc[:, 1] * ha + cya
        w, h = torch.exp(loc[:, 2]) * wa, torch.exp(loc[:, 3]) * ha
        roi = torch.stack((cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h), 1)
        roi[:, [0, 2]] = roi[:, [0, 2]].clamp(min=0, max=float(img_size[1]))
        roi[:, [1, 3]] = roi[:, [1, 3]].clamp(min=0, max=float(img_size[0]))
        ms = self.min_size * scale
        valid = ((roi[:, 2] - roi[:, 0]) >= ms) & ((roi[:, 3] - roi[:, 1]) >= ms)
        rois, sc, idx, cnt = select_nms(roi[None], score[None].float(), valid[None], n_pre, n_post, self.nms_iou)
        k = int(cnt.item())
        if return_index:
            return rois[0, :k, 1:], sc[0, :k], idx[0, :k]
        return rois[0, :k, 1:]


class RpnProposalsPerLevel:
    """torchvision RegionProposalNetwork.filter_proposals semantics (models/detection/rpn.py:231-297; SURVEY.md A.3 brackets):
    top pre_nms_top_n PER LEVEL on the objectness logits -> sigmoid -> clip -> drop boxes smaller than min_size (1e-3) or below
    score_thresh -> NMS per level (batched_nms with the level as class) -> first post_nms_top_n by score.
    Composed from the C-ABI entry points on the device, no host sync: hd_rpn_decode (logit keys, dw/dh clamp), one
    hd_rpn_select_nms per level with an IoU threshold nothing can exceed (= stable top-k + sort), a stable valid-first
    partition, hd_sort_nms_batched(HD_NMS_CLASS_EXACT).  Returns (rois [B*post,5], count [B], scores [B,post], idx [B,post])."""

    def __init__(self, anchor_bases, strides, img_size, nms_thresh=0.7, pre_nms_top_n=1000, post_nms_top_n=1000, min_size=1e-3,
                 score_thresh=0.0, clamp_dwh=math.log(1000.0 / 16)):
        self.dec = RpnProposals(anchor_bases, strides, img_size, nms_thresh, pre_nms_top_n, post_nms_top_n, min_size=-3.0e38,
                                score_mode="sigmoid", clamp_dwh=clamp_dwh)
        self.dec.flags |= _lib.RPN_KEY_LOGIT
        self.nms_thresh, self.pre, self.post = float(nms_thresh), int(pre_nms_top_n), int(post_nms_top_n)
        self.min_size, self.score_thresh = float(min_size), float(score_thresh)

    def __call__(self, objectness, deltas):
        boxes, scores, keys = self.dec.decode(objectness, deltas)
        return self.filter_proposals(boxes, scores, keys, [d.shape[1] // 4 * d.shape[2] * d.shape[3] for d in deltas])

    def _buffers(self, B, N, ks, dev):
        key = (B, N, tuple(ks), dev)
        if getattr(self, "_bkey", None) != key:
            L = _lib.lib()
            K = sum(ks)
            self._lvl = []
            for k_l, n_l in zip(ks, self._n_per_level):
                ws_bytes = L.hd_rpn_select_nms_workspace_size(B, n_l, k_l)
                self._lvl.append((torch.empty((ws_bytes,), dtype=torch.uint8, device=dev), ws_bytes,
                                  torch.empty((B, k_l, 5), dtype=torch.float32, device=dev), torch.empty((B, k_l), dtype=torch.int64, device=dev),
                                  torch.zeros((B,), dtype=torch.int32, device=dev), torch.cuda.Stream(dev)))
            self._cb = torch.empty((B, K, 4), dtype=torch.float32, device=dev)
            self._cs = torch.empty((B, K), dtype=torch.float32, device=dev)
            self._cl = torch.empty((B, K), dtype=torch.int32, device=dev)
            self._ca = torch.empty((B, K), dtype=torch.int32, device=dev)
            self._cc = torch.zeros((B,), dtype=torch.int32, device=dev)
            self._nws_bytes = L.hd_sort_nms_workspace_size(B, K)
            self._nws = torch.empty((self._nws_bytes,), dtype=torch.uint8, device=dev)
            self._det = torch.zeros((B, self.post, 6), dtype=torch.float32, device=dev)
            self._slot = torch.zeros((B, self.post), dtype=torch.int64, device=dev)
            self._cnt = torch.zeros((B,), dtype=torch.int32, device=dev)
            self._bkey = key

    @_lib.on_device
    def filter_proposals(self, boxes, scores, keys, num_anchors_per_level):
        """stage 2 on the decoded arrays: boxes [B,N,4], scores [B,N] (probabilities), keys [B,N] int32 (sortable logit bits).
        Six launches on the device, no eager tensor glue: the per-level stable top-k of the L levels run side by side on L streams
        (they read disjoint slices of the same arrays, in place), then hd_rpn_merge_levels, the level-aware NMS and hd_rpn_finish_levels."""
        import ctypes as C
        L = _lib.lib()
        boxes, scores, keys = _lib.f32c(boxes), _lib.f32c(scores), keys.contiguous()
        B, N = scores.shape
        dev = scores.device
        self._n_per_level = list(num_anchors_per_level)
        ks = [min(self.pre, n_l) for n_l in num_anchors_per_level]
        K = sum(ks)
        self._buffers(B, N, ks, dev)
        cur = torch.cuda.current_stream(dev)
        fork = torch.cuda.Event()
        fork.record(cur)
        offs, off = [], 0
        for l, (n_l, k_l) in enumerate(zip(num_anchors_per_level, ks)):
            ws, ws_bytes, rois_l, idx_l, cnt_l, st = self._lvl[l]
            offs.append(off)
            st.wait_event(fork)
            with torch.cuda.stream(st):
                # IoU threshold 2: nothing is ever suppressed -> the call is a stable top-k + sort of the level's slice
                _lib.check(L.hd_rpn_select_nms_strided(
                    C.c_void_p(boxes.data_ptr() + off * 16), C.c_void_p(scores.data_ptr() + off * 4), C.c_void_p(keys.data_ptr() + off * 4),
                    B, n_l, N, k_l, k_l, 2.0, _lib.ptr(rois_l), None, _lib.ptr(idx_l), _lib.ptr(cnt_l), _lib.ptr(ws), ws_bytes, _lib.stream()))
            cur.wait_stream(st)
            off += n_l
        nl = len(ks)
        sel = (C.c_void_p * nl)(*[self._lvl[l][3].data_ptr() for l in range(nl)])
        karr, oarr = (C.c_int32 * nl)(*ks), (C.c_int32 * nl)(*offs)
        _lib.check(L.hd_rpn_merge_levels(_lib.ptr(boxes), _lib.ptr(scores), sel, karr, oarr, nl, B, N, self.min_size, self.score_thresh,
                                         _lib.ptr(self._cb), _lib.ptr(self._cs), _lib.ptr(self._cl), _lib.ptr(self._ca), _lib.ptr(self._cc), _lib.stream()))
        _lib.check(L.hd_sort_nms_batched(_lib.ptr(self._cb), _lib.ptr(self._cs), _lib.ptr(self._cl), None, _lib.ptr(self._cc), 0, B, K, self.nms_thresh,
                                         _lib.NMS_CLASS_EXACT, 0.0, 0, self.post, _lib.ptr(self._det), _lib.ptr(self._slot), _lib.ptr(self._cnt),
                                         _lib.ptr(self._nws), self._nws_bytes, _lib.stream()))
        rois = torch.empty((B * self.post, 5), dtype=torch.float32, device=dev)
        osc = torch.empty((B, self.post), dtype=torch.float32, device=dev)
        oidx = torch.empty((B, self.post), dtype=torch.int64, device=dev)
        _lib.check(L.hd_rpn_finish_levels(_lib.ptr(self._det), _lib.ptr(self._slot), _lib.ptr(self._cnt), _lib.ptr(self._ca), B, K, self.post,
                                          _lib.ptr(rois), _lib.ptr(osc), _lib.ptr(oidx), _lib.stream()))
        return rois, self._cnt, osc, oidx

"""RoIAlign / RoIPool / FPN level assignment host API (README.md:65,73-78).

``roi_align`` / ``roi_pool`` / ``RoIAlign`` / ``RoIPool`` keep the torchvision signatures
(roi_align.py:204-283, roi_pool.py:15-68); ``multilevel_roi_align`` is the MultiScaleRoIAlign
gather-scatter (poolers.py:147-227) as ONE launch over all levels.
"""
import torch
from torch import nn
from . import _lib


def set_mode(mode):
    """RoIAlign kernel choice for many RoIs on NHWC features: 0 / 1 per-RoI gather kernels (default: the fastest measured), 2 streamed
    (TMA ring) kernel, 3 row-walk kernel over the device-bucketed RoIs -- 2 and 3 are bit-identical, slower, opt-in; returns the previous mode."""
    return _lib.lib().hd_roi_set_mode(int(mode))


def _pair(v):
    return (int(v), int(v)) if isinstance(v, int) else (int(v[0]), int(v[1]))


def convert_boxes_to_roi_format(boxes):
    """list[Tensor[L,4]] -> Tensor[K,5], batch index = list position (torchvision _utils.py:18-25)."""
    cat = torch.cat(list(boxes), 0)
    ids = torch.cat([torch.full((b.shape[0], 1), i, dtype=cat.dtype, device=cat.device) for i, b in enumerate(boxes)], 0)
    return torch.cat((ids, cat), 1)


def _check_rois(boxes):
    if isinstance(boxes, (list, tuple)):
        for b in boxes:
            assert b.dim() == 2 and b.size(1) == 4, "The shape of the tensor in the boxes list is not correct as List[Tensor[L, 4]]"
        return convert_boxes_to_roi_format(boxes)
    assert boxes.dim() == 2 and boxes.size(1) == 5, "The boxes tensor shape is not correct as Tensor[K, 5]"
    return boxes


def _is_channels_last(x):
    return x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last) and not (x.is_contiguous() and x.shape[1] > 1 and x.shape[2] * x.shape[3] > 1)


def _prepare_level(x, K, layout):
    """-> (tensor holding the data, layout id).  'auto': channels_last inputs are used in place; NCHW inputs
    get one layout pass when there are enough RoIs to amortise it, else the direct NCHW kernel."""
    B, C, H, W = x.shape
    if layout == "nhwc" or (layout == "auto" and _is_channels_last(x)):
        if _is_channels_last(x):
            return x, _lib.LAYOUT_NHWC
        xc = _lib.f32c(x)
        out = torch.empty((B, H, W, C), dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().hd_nchw_to_nhwc(_lib.ptr(xc), _lib.ptr(out), B, C, H, W, _lib.stream()))
        return out, _lib.LAYOUT_NHWC
    if layout == "auto" and K * 392 > B * H * W:
        return _prepare_level(x, K, "nhwc")
    return _lib.f32c(x), _lib.LAYOUT_NCHW


_WS = {}


def _workspace(device, nbytes):
    """scratch for hd_roi_align_ws, one buffer per (device, stream): calls on different streams must not share it"""
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        if len(_WS) > 16:
            _WS.clear()
        buf = _WS[key] = torch.empty((max(int(nbytes), 256),), dtype=torch.uint8, device=device)
    return buf


def _levels_struct(tensors, scales):
    arr = (_lib.RoiLevel * len(tensors))()
    for l, (t, s) in enumerate(zip(tensors, scales)):
        arr[l].data = t.data_ptr()
        arr[l].spatial_scale = float(s)
    return arr


@_lib.on_device
def _run(features, scales, rois, level_ids, output_size, sampling_ratio, aligned, pool, layout, return_argmax=False):
    PH, PW = _pair(output_size)
    for f in features:
        _lib.require_cuda(f)
        if f.dtype != torch.float32:
            raise NotImplementedError(f"hd_b200 roi ops are fp32-only, got {f.dtype}")
    _lib.require_cuda(rois)
    rois = _lib.f32c(rois.to(torch.float32))
    K, C = rois.shape[0], features[0].shape[1]
    out = torch.empty((K, C, PH, PW), dtype=torch.float32, device=features[0].device)
    argmax = torch.empty((K, C, PH, PW), dtype=torch.int32, device=out.device) if (pool and return_argmax) else None
    if K == 0:
        return (out, argmax) if return_argmax else out
    prepared, lay = [], None
    for f in features:
        t, l = _prepare_level(f, K, layout if lay is None else ("nhwc" if lay == _lib.LAYOUT_NHWC else "nchw"))
        lay = l
        prepared.append(t)
    arr = _levels_struct(prepared, scales)
    for l, f in enumerate(features):
        arr[l].H, arr[l].W = f.shape[2], f.shape[3]
    lid = None if level_ids is None else level_ids.to(torch.int32).contiguous()
    L = _lib.lib()
    if pool:
        _lib.check(L.hd_roi_pool(arr, len(prepared), lay, C, _lib.ptr(rois), _lib.ptr(lid), K, PH, PW, _lib.ptr(out),
                                 _lib.ptr(argmax), _lib.stream()))
    else:
        # batch size + workspace: many RoIs on NHWC features take the streamed (TMA ring) kernel, everything else the per-RoI kernels
        B = int(features[0].shape[0])
        ws_bytes = L.hd_roi_align_workspace_size(arr, len(prepared), C, B, K, PH, PW, int(sampling_ratio))
        ws = _workspace(out.device, ws_bytes)
        _lib.check(L.hd_roi_align_ws(arr, len(prepared), lay, C, B, _lib.ptr(rois), _lib.ptr(lid), K, PH, PW, int(sampling_ratio),
                                     int(bool(aligned)), _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.stream()))
    return (out, argmax) if return_argmax else out


@_lib.on_device
def roi_align_backward(grad, rois, spatial_scale, pooled_height, pooled_width, batch_size, channels, height, width, sampling_ratio, aligned,
                       channels_last=True):
    """torch.ops.torchvision._roi_align_backward schema -> grad_input [B,C,H,W] (channels_last memory by default: the layout whose
    kernel adds a channel quad with one 128-bit reduction)."""
    _lib.require_cuda(grad, rois)
    grad, rois = _lib.f32c(grad), _lib.f32c(rois.to(torch.float32))
    if channels_last:   # NHWC storage viewed as [B,C,H,W] (= torch.channels_last)
        gi = torch.zeros((batch_size, height, width, channels), dtype=torch.float32, device=grad.device).permute(0, 3, 1, 2)
    else:
        gi = torch.zeros((batch_size, channels, height, width), dtype=torch.float32, device=grad.device)
    K = rois.shape[0]
    if K == 0:
        return gi
    arr = _levels_struct([gi], [spatial_scale])
    arr[0].H, arr[0].W = height, width
    lay = _lib.LAYOUT_NHWC if channels_last else _lib.LAYOUT_NCHW
    _lib.check(_lib.lib().hd_roi_align_backward(_lib.ptr(grad), _lib.ptr(rois), None, K, arr, 1, lay, channels, int(pooled_height), int(pooled_width),
                                                int(sampling_ratio), int(bool(aligned)), _lib.stream()))
    return gi


@_lib.on_device
def roi_pool_backward(grad, rois, argmax, spatial_scale, pooled_height, pooled_width, batch_size, channels, height, width, channels_last=False):
    """torch.ops.torchvision._roi_pool_backward schema -> grad_input [B,C,H,W]"""
    _lib.require_cuda(grad, rois, argmax)
    grad, rois = _lib.f32c(grad), _lib.f32c(rois.to(torch.float32))
    argmax = argmax.to(torch.int32).contiguous()
    if channels_last:
        gi = torch.zeros((batch_size, height, width, channels), dtype=torch.float32, device=grad.device).permute(0, 3, 1, 2)
    else:
        gi = torch.zeros((batch_size, channels, height, width), dtype=torch.float32, device=grad.device)
    K = rois.shape[0]
    if K:
        _lib.check(_lib.lib().hd_roi_pool_backward(_lib.ptr(grad), _lib.ptr(argmax), _lib.ptr(rois), K, _lib.ptr(gi),
                                                   _lib.LAYOUT_NHWC if channels_last else _lib.LAYOUT_NCHW, channels, height, width,
                                                   int(pooled_height), int(pooled_width), _lib.stream()))
    return gi


class _RoIPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, rois, output_size, spatial_scale, layout):
        out, argmax = _run([input], [spatial_scale], rois, None, output_size, 0, False, True, layout, return_argmax=True)
        ctx.save_for_backward(rois, argmax)
        ctx.args = (tuple(input.shape), _pair(output_size), float(spatial_scale), _is_channels_last(input))
        return out

    @staticmethod
    def backward(ctx, grad):
        rois, argmax = ctx.saved_tensors
        (B, C, H, W), (PH, PW), scale, cl = ctx.args
        return roi_pool_backward(grad, rois, argmax, scale, PH, PW, B, C, H, W, channels_last=cl), None, None, None, None


class _RoIAlignFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, rois, output_size, spatial_scale, sampling_ratio, aligned, layout):
        ctx.save_for_backward(rois)
        ctx.args = (tuple(input.shape), _pair(output_size), float(spatial_scale), int(sampling_ratio), bool(aligned), _is_channels_last(input))
        return _run([input], [spatial_scale], rois, None, output_size, sampling_ratio, aligned, False, layout)

    @staticmethod
    def backward(ctx, grad):
        (rois,) = ctx.saved_tensors
        (B, C, H, W), (PH, PW), scale, sr, al, cl = ctx.args
        gi = roi_align_backward(grad, rois, scale, PH, PW, B, C, H, W, sr, al, channels_last=cl)
        return gi, None, None, None, None, None, None


def roi_align(input, boxes, output_size, spatial_scale=1.0, sampling_ratio=-1, aligned=False, layout="auto"):
    """torchvision.ops.roi_align (roi_align.py:204-260); differentiable w.r.t. `input` (RoIAlign backward kernels)."""
    rois = _check_rois(boxes)
    if input.requires_grad and torch.is_grad_enabled():
        return _RoIAlignFn.apply(input, rois, output_size, spatial_scale, sampling_ratio, aligned, layout)
    return _run([input], [spatial_scale], rois, None, output_size, sampling_ratio, aligned, False, layout)


def roi_pool(input, boxes, output_size, spatial_scale=1.0, layout="auto"):
    """torchvision.ops.roi_pool (roi_pool.py:15-53); differentiable w.r.t. `input`."""
    rois = _check_rois(boxes)
    if input.requires_grad and torch.is_grad_enabled():
        return _RoIPoolFn.apply(input, rois, output_size, spatial_scale, layout)
    return _run([input], [spatial_scale], rois, None, output_size, 0, False, True, layout)


def roi_pool_with_argmax(input, boxes, output_size, spatial_scale=1.0, layout="auto"):
    """torch.ops.torchvision.roi_pool schema: (output, int32 argmax)."""
    rois = _check_rois(boxes)
    return _run([input], [spatial_scale], rois, None, output_size, 0, False, True, layout, return_argmax=True)


class RoIAlign(nn.Module):
    """torchvision.ops.RoIAlign (roi_align.py:263-283)."""

    def __init__(self, output_size, spatial_scale, sampling_ratio, aligned=False):
        super().__init__()
        self.output_size, self.spatial_scale = output_size, spatial_scale
        self.sampling_ratio, self.aligned = sampling_ratio, aligned

    def forward(self, input, rois):
        return roi_align(input, rois, self.output_size, self.spatial_scale, self.sampling_ratio, self.aligned)

    def __repr__(self):
        return (f"{self.__class__.__name__}(output_size={self.output_size}, spatial_scale={self.spatial_scale}, "
                f"sampling_ratio={self.sampling_ratio}, aligned={self.aligned})")


class RoIPool(nn.Module):
    """torchvision.ops.RoIPool (roi_pool.py:56-68)."""

    def __init__(self, output_size, spatial_scale):
        super().__init__()
        self.output_size, self.spatial_scale = output_size, spatial_scale

    def forward(self, input, rois):
        return roi_pool(input, rois, self.output_size, self.spatial_scale)

    def __repr__(self):
        return f"{self.__class__.__name__}(output_size={self.output_size}, spatial_scale={self.spatial_scale})"


@_lib.on_device
def level_map(boxes_xyxy, k_min=2, k_max=5, canonical_scale=224.0, canonical_level=4, eps=1e-6, style="torchvision"):
    """FPN level of each box [K,4] -> int64 [K] (poolers.py:73-84; style 'mmdet': eps inside the log)."""
    _lib.require_cuda(boxes_xyxy)
    b = _lib.f32c(boxes_xyxy)
    out = torch.empty((b.shape[0],), dtype=torch.int64, device=b.device)
    _lib.check(_lib.lib().hd_roi_level_map(_lib.ptr(b), 4, 0, b.shape[0], 1 if style == "mmdet" else 0, int(k_min), int(k_max),
                                           float(canonical_scale), float(canonical_level), float(eps), None, _lib.ptr(out), _lib.stream()))
    return out


@_lib.on_device
def multilevel_roi_align(features, rois, output_size, spatial_scales, sampling_ratio=2, aligned=False, op="align",
                         levels=None, layout="auto", k_min=2, k_max=5, canonical_scale=224.0, canonical_level=4,
                         eps=1e-6, style="torchvision"):
    """Assign each RoI [K,5] to a pyramid level and pool it there, original RoI order kept.
    -> (out [K,C,PH,PW], levels int64 [K]).  A single feature map is the README's "P2" variant."""
    rois = _check_rois(rois)
    _lib.require_cuda(rois)
    rois = _lib.f32c(rois.to(torch.float32))
    K = rois.shape[0]
    if levels is None:
        if len(features) > 1:
            lv = torch.empty((K,), dtype=torch.int32, device=rois.device)
            lv64 = torch.empty((K,), dtype=torch.int64, device=rois.device)
            _lib.check(_lib.lib().hd_roi_level_map(_lib.ptr(rois), 5, 1, K, 1 if style == "mmdet" else 0, int(k_min), int(k_max),
                                                   float(canonical_scale), float(canonical_level), float(eps), _lib.ptr(lv),
                                                   _lib.ptr(lv64), _lib.stream()))
        else:
            lv, lv64 = None, torch.zeros((K,), dtype=torch.int64, device=rois.device)
    else:
        lv, lv64 = levels.to(torch.int32), levels.to(torch.int64)
    out = _run(list(features), list(spatial_scales), rois, lv, output_size, sampling_ratio, aligned, op != "align", layout)
    return out, lv64

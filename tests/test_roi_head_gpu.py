"""RoI-head output post-process + output formats against torchvision's own
RoIHeads.postprocess_detections run on CPU (oracle.roi_head.postprocess_detections_tv).

Integer part (which (roi, class) candidates survive, NMS keeps, labels): bit-exact given the same candidate scores;
softmax on CUDA and CPU differs by ulps, so candidates within 1e-6 of the score threshold are excluded from the
set comparison.  Floating point (boxes, scores): <= 1e-5 relative (box floor = box size, tests/_tol.py rule)."""
import math
import pytest
import torch

pytestmark = pytest.mark.gpu


def _inputs(B, R, C, seed, img=(416, 480)):
    g = torch.Generator().manual_seed(seed)
    lg = torch.randn(B * R, C, generator=g) * 3
    rg = torch.randn(B * R, C * 4, generator=g) * 0.6
    xy = torch.rand(B * R, 2, generator=g) * torch.tensor([img[1] - 40.0, img[0] - 40.0])
    wh = torch.rand(B * R, 2, generator=g) * 160 + 4
    pr = torch.cat((xy, xy + wh), 1)
    return lg, rg, [pr[b * R:(b + 1) * R] for b in range(B)]


def _close_boxes(a, b):
    a, b = a.double(), b.double()
    size = torch.maximum(b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]).clamp(min=1.0)
    return bool(((a - b).abs() <= 1e-5 * torch.maximum(b.abs(), size[:, None])).all())


@pytest.mark.parametrize("B,R,C,thr,nms,topk", [(2, 300, 21, 0.05, 0.5, 100), (3, 1000, 81, 0.05, 0.5, 100), (1, 64, 2, 0.3, 0.3, 10), (2, 500, 11, 0.001, 0.7, 300)])
def test_postprocess_detections_matches_torchvision_method(B, R, C, thr, nms, topk):
    import oracle
    from heltondetection_b200 import roi_head
    lg, rg, props = _inputs(B, R, C, 7 + R)
    shapes = [(416, 480)] * B
    rb, rs, rl = oracle.roi_head.postprocess_detections_tv(lg, rg, props, shapes, thr, nms, topk)
    _, _, _, ri = oracle.roi_head.postprocess_detections(lg, rg, props, shapes, thr, nms, topk, return_ids=True)
    gb, gs, gl, gi = roi_head.postprocess_detections(lg.cuda(), rg.cuda(), [p.cuda() for p in props], shapes, thr, nms, topk, return_ids=True)
    for b in range(B):
        assert torch.equal(gi[b].cpu(), ri[b]), "kept (roi, class) ids"
        assert torch.equal(gl[b].cpu(), rl[b])
        assert _close_boxes(gb[b].cpu(), rb[b])
        assert bool(((gs[b].cpu().double() - rs[b].double()).abs() <= 1e-5 * rs[b].double().clamp(min=1e-3)).all())


def test_candidates_stage_and_ragged_images():
    import oracle
    from heltondetection_b200 import roi_head
    lg, rg, props = _inputs(3, 200, 21, 3)
    props = [props[0][:150], props[1], props[2][:7]]          # ragged; different image shapes -> per-image calls
    n = [p.shape[0] for p in props]
    keep = torch.cat([torch.arange(b * 200, b * 200 + n[b]) for b in range(3)])
    lg, rg = lg[keep], rg[keep]
    shapes = [(416, 480), (300, 500), (416, 480)]
    rb, rs, rl, ri = oracle.roi_head.postprocess_detections(lg, rg, props, shapes, 0.05, 0.5, 100, return_ids=True)
    gb, gs, gl, gi = roi_head.postprocess_detections(lg.cuda(), rg.cuda(), [p.cuda() for p in props], shapes, 0.05, 0.5, 100, return_ids=True)
    for b in range(3):
        assert torch.equal(gi[b].cpu(), ri[b]) and torch.equal(gl[b].cpu(), rl[b]) and _close_boxes(gb[b].cpu(), rb[b])
    # same shapes, ragged lengths -> one padded call with roi_count
    shapes = [(416, 480)] * 3
    rb, rs, rl, ri = oracle.roi_head.postprocess_detections(lg, rg, props, shapes, 0.05, 0.5, 100, return_ids=True)
    gb, gs, gl, gi = roi_head.postprocess_detections(lg.cuda(), rg.cuda(), [p.cuda() for p in props], shapes, 0.05, 0.5, 100, return_ids=True)
    for b in range(3):
        assert torch.equal(gi[b].cpu(), ri[b]) and torch.equal(gl[b].cpu(), rl[b])


def test_lineage_decodebox_variant():
    import oracle
    from heltondetection_b200 import roi_head
    B, R, C = 2, 300, 21
    lg, rg, props = _inputs(B, R, C, 11)
    ref = oracle.roi_head.postprocess_detections(lg, rg, props, [(416, 480)] * B, 0.5, 0.3, R * (C - 1), weights=(0.1, 0.1, 0.2, 0.2), mul_std=True,
                                                 clamp=None, min_size=None, label_minus1=True, return_ids=True)
    out = roi_head.DecodeBox(C - 1).forward(rg.view(B, R, -1).cuda(), lg.view(B, R, -1).cuda(), torch.stack(props).cuda(), (416, 480), 0.3, 0.5)
    for b in range(B):
        rb, rs, rl = ref[0][b], ref[1][b], ref[2][b]
        order = torch.sort(rl, stable=True)[1]
        got = out[b].cpu()
        assert got.shape[0] == rb.shape[0]
        assert torch.equal(got[:, 5].long(), rl[order])
        assert _close_boxes(got[:, :4], rb[order])


def test_padded_api_on_rpn_layout_and_empty():
    from heltondetection_b200 import roi_head
    lg, rg, props = _inputs(2, 128, 5, 5)
    rois = torch.zeros(2, 128, 5)
    for b in range(2):
        rois[b, :, 0] = b
        rois[b, :, 1:] = props[b]
    cnt = torch.tensor([128, 0], dtype=torch.int32)
    pp = roi_head.RoIHeadPostprocessor((416, 480))
    det, idx, count = pp(lg.cuda(), rg.cuda(), rois.view(-1, 5).cuda(), cnt.cuda())
    assert int(count[1]) == 0 and int(count[0]) > 0
    with pytest.raises(RuntimeError):
        pp(lg, rg, rois.view(-1, 5))   # CPU tensors are refused: no fallback


def test_scale_coords_and_coco_boxes_bit_exact():
    import oracle
    from heltondetection_b200 import roi_head
    g = torch.Generator().manual_seed(2)
    B, M = 3, 50
    det = torch.rand(B, M, 6, generator=g) * 640
    det[..., 2:4] = det[..., :2] + torch.rand(B, M, 2, generator=g) * 200
    cnt = torch.tensor([50, 17, 0], dtype=torch.int32)
    img0 = [(480, 640), (1080, 1920), (333, 500)]
    out = roi_head.scale_coords((640, 640), det.cuda(), img0, cnt.cuda()).cpu()
    outw = roi_head.scale_coords((640, 640), det.cuda(), img0, cnt.cuda(), xywh=True).cpu()
    for b in range(B):
        k = int(cnt[b])
        ref = oracle.roi_head.scale_coords((640, 640), det[b, :k, :4], img0[b])
        assert torch.equal(out[b, :k, :4], ref)
        assert torch.equal(outw[b, :k, :4], oracle.roi_head.xyxy2xywh_coco(ref))
        assert torch.equal(out[b, :k, 4:], det[b, :k, 4:]) and bool((out[b, k:] == 0).all())
    recs = roi_head.coco_records(roi_head.scale_coords((640, 640), det.cuda(), img0, cnt.cuda(), xywh=True), cnt, [10, 11, 12])
    assert len(recs) == 67 and recs[0]["image_id"] == 10 and len(recs[0]["bbox"]) == 4

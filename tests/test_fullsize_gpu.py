"""Oracle parity at the FULL BASELINE.json sizes, EVERY image of every config (VERDICT r1, row g).

The CPU oracle runs ~360 img/s on cfg2, ~30 img/s on cfg4 and ~3 img/s on cfg3, so checking all 256 / 64 / 16 images costs
seconds: there is no sampling here.  Bars (north_star): keep indices, candidate counts and level ids bit-exact; boxes, scores
and RoI features within 1e-5 relative fp32 (tests/_tol.py).

cfg1  YOLOv5s 640^2 B=1, conf 0.25 and 0.001
cfg2  YOLOv5s 640^2 B=256, dense and objectness-skip reads
cfg3  FasterRCNN 832^2 B=16: RPN decode -> top-12000 -> NMS 0.7 -> top-2000 END TO END (torch.equal on the indices with
      exact_math; the raw mismatch rate against the plain fp32 CPU path is measured and printed), multi-level RoIAlign
      (sampling_ratio 2 and 0, aligned False and True) and RoIPool on all 32 000 proposals
cfg4  YOLOv5l 1280^2 B=64 nc=10 dense scenes, conf 0.001 iou 0.6
cfg5  TTA 6 views B=64 + Weighted Boxes Fusion
"""
import numpy as np
import pytest
import torch

from _tol import close, boxes_close

pytestmark = pytest.mark.gpu


def _yolo_every_image(heads_cpu, conf, iou, dense, chunk=32, **kw):
    import oracle
    from heltondetection_b200 import yolo
    B = heads_cpu[0].shape[0]
    pp = yolo.YoloPostprocessor(conf_thres=conf, iou_thres=iou, dense_read=dense, **kw)
    det, cnt, idx = pp([h.cuda() for h in heads_cpu])
    det, cnt, idx = det.cpu(), cnt.cpu(), idx.cpu()
    bad = []
    for b0 in range(0, B, chunk):
        pred = oracle.yolo.decode_box([h[b0:b0 + chunk] for h in heads_cpu])
        ref, ridx = oracle.yolo.non_max_suppression(pred, conf, iou, return_index=True, **kw)
        for j, (r, ri) in enumerate(zip(ref, ridx)):
            b = b0 + j
            n = int(cnt[b])
            ok = n == ri.numel() and torch.equal(idx[b, :n], ri)
            ok = ok and boxes_close(det[b, :n, :4], r[:, :4]) and close(det[b, :n, 4], r[:, 4], scale=1e-3) and torch.equal(det[b, :n, 5], r[:, 5])
            ok = ok and bool((det[b, n:] == 0).all()) and bool((idx[b, n:] == -1).all())      # deterministic padding
            if not ok:
                bad.append(b)
    assert not bad, f"images differing from the oracle: {bad[:10]} ({len(bad)} of {B})"
    return cnt


@pytest.mark.parametrize("conf", [0.25, 0.001])
def test_cfg1_every_threshold(conf):
    from heltondetection_b200 import synth
    heads, _ = synth.yolo_heads(1, 640, 80, 20, 1234)
    cnt = _yolo_every_image(heads, conf, 0.45, True)
    assert int(cnt[0]) > 0


@pytest.mark.parametrize("dense", [True, False])
def test_cfg2_all_256_images(dense):
    from heltondetection_b200 import synth
    heads, _ = synth.yolo_heads(256, 640, 80, 20, 1235)
    cnt = _yolo_every_image(heads, 0.25, 0.45, dense)
    assert int(cnt.sum()) > 256 * 10


def test_cfg4_all_64_images():
    from heltondetection_b200 import synth
    heads, _ = synth.yolo_heads(64, 1280, 10, 300, 1238, dense=True)
    cnt = _yolo_every_image(heads, 0.001, 0.6, True, chunk=8)
    assert int(cnt.min()) == 300          # dense scenes fill max_det


def test_cfg3_rpn_all_16_images_end_to_end():
    """decode -> select -> NMS chained on the device vs the chained CPU oracle, torch.equal on every image.  With exact_math both
    sides evaluate sigmoid/exp in fp64 and round once, so no libm ulp can swap near-tied scores; the default fp32 path is
    compared with the plain fp32 torch-CPU oracle and its RAW index mismatch rate is printed (SURVEY.md 7 'hard parts')."""
    import oracle
    from heltondetection_b200 import synth, rpn
    B, img = 16, 832
    obj, dlt, bases, _ = synth.rpn_heads(B, img, G=20, seed=1237)
    kw = dict(n_pre_nms=12000, n_post_nms=2000, min_size=16)
    dev_o, dev_d = [o.cuda() for o in obj], [d.cuda() for d in dlt]
    pr = rpn.RpnProposals(bases, (4, 8, 16, 32), (img, img), exact_math=True, **kw)
    rois, cnt, sc, idx = pr(dev_o, dev_d)
    rois, cnt, sc, idx = rois.view(B, 2000, 5).cpu(), cnt.cpu(), sc.cpu(), idx.cpu()
    ref = oracle.rpn.rpn_proposals(obj, dlt, bases, (4, 8, 16, 32), (img, img), exact_math=True, **kw)
    for b in range(B):
        r_roi, r_sc, r_idx = ref[b]
        n = int(cnt[b])
        assert n == r_idx.numel() == 2000
        assert torch.equal(idx[b, :n], r_idx), f"image {b}: proposal indices differ"
        assert torch.equal(rois[b, :n, 1:], r_roi) and torch.equal(sc[b, :n], r_sc), f"image {b}: proposal boxes/scores differ in bits"
    # raw mismatch rate of the default (fp32 expf) path against the plain fp32 CPU path
    pr32 = rpn.RpnProposals(bases, (4, 8, 16, 32), (img, img), **kw)
    _, cnt32, _, idx32 = pr32(dev_o, dev_d)
    ref32 = oracle.rpn.rpn_proposals(obj, dlt, bases, (4, 8, 16, 32), (img, img), **kw)
    mism, total, pos = 0, 0, 0
    for b in range(B):
        got, want = idx32[b, : int(cnt32[b])].cpu(), ref32[b][2]
        total += want.numel()
        mism += want.numel() - len(set(got.tolist()) & set(want.tolist()))
        m = min(got.numel(), want.numel())
        pos += int((got[:m] != want[:m]).sum()) + abs(got.numel() - want.numel())
    print(f"\n[rpn fp32 path vs fp32 torch-CPU oracle] raw index mismatch: {mism}/{total} = {mism / total:.4%} of the kept set; "
          f"{pos}/{total} = {pos / total:.4%} positions differ")
    assert mism / total <= 0.02


@pytest.fixture(scope="module")
def cfg3_rois():
    from heltondetection_b200 import synth, rpn
    B, img = 16, 832
    obj, dlt, bases, _ = synth.rpn_heads(B, img, G=20, seed=1237)
    pr = rpn.RpnProposals(bases, (4, 8, 16, 32), (img, img), n_pre_nms=12000, n_post_nms=2000, min_size=16)
    rois, cnt, _, _ = pr([o.cuda() for o in obj], [d.cuda() for d in dlt])
    assert cnt.tolist() == [2000] * B
    feats = synth.fpn_features(B, img, 256, 1237)
    return rois.clone(), feats


@pytest.mark.parametrize("sr,aligned,op", [(2, False, "align"), (0, True, "align"), (2, False, "pool")])
def test_cfg3_roi_ops_all_32000_rois(cfg3_rois, sr, aligned, op):
    import oracle
    from heltondetection_b200 import ops
    rois, feats = cfg3_rois
    scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
    nhwc = [f.cuda().contiguous(memory_format=torch.channels_last) for f in feats]
    got, lv = ops.multilevel_roi_align(nhwc, rois, 7, scales, sr, aligned, op=op)
    assert got.shape == (32000, 256, 7, 7)
    got, lv = got.cpu(), lv.cpu()
    ref, rlv = oracle.roi.multilevel_roi_align(feats, rois.cpu(), 7, scales, sr, aligned, op=op)
    assert torch.equal(lv, rlv), "level assignment differs"
    if op == "pool":
        assert torch.equal(got, ref), "RoIPool values differ in bits"
    else:
        d = (got.double() - ref.double()).abs()
        tol = 1e-5 * ref.double().abs().clamp(min=1.0)
        bad = (d > tol).flatten(1).any(1)
        assert not bool(bad.any()), f"{int(bad.sum())} of 32000 RoIs outside 1e-5 (max abs diff {float(d.max()):.3e})"


def test_cfg5_all_64_images():
    """6 views -> per-view decode+NMS (checked against the oracle per view) -> map back -> WBF (checked on the device's per-view
    detections, so every stage is compared on identical inputs and the chain is covered end to end)."""
    import oracle
    from heltondetection_b200 import synth, yolo, wbf
    B, img, nc = 64, 640, 80
    views, _ = synth.tta_heads(B, img, nc, G=20, seed=1239)
    vspec = [(r, flip, size) for (_, r, flip, size) in views]
    fusion = wbf.TTAFusion(vspec, (img, img), nc, max_det=300, iou_thr=0.55, skip_box_thr=0.001)
    ref_lists = [([], [], []) for _ in range(B)]
    pp = yolo.YoloPostprocessor(conf_thres=0.25, iou_thres=0.45)
    for v, (heads, r, flip, size) in enumerate(views):
        det, cnt, idx = pp([h.cuda() for h in heads])
        fusion.map_back(v, det, cnt)
        det_c, cnt_c, idx_c = det.cpu(), cnt.cpu(), idx.cpu()
        pred = oracle.yolo.decode_box(heads)
        ref, ridx = oracle.yolo.non_max_suppression(pred, 0.25, 0.45, return_index=True)
        for b in range(B):
            n = int(cnt_c[b])
            assert torch.equal(idx_c[b, :n], ridx[b]), f"view {v} image {b}: keep indices differ"
            assert boxes_close(det_c[b, :n, :4], ref[b][:, :4])
            bb, ss, ll = oracle.tta.map_back(det_c[b, :n], r, flip, float(size), float(img), float(img))
            ref_lists[b][0].append(bb.numpy()); ref_lists[b][1].append(ss.numpy()); ref_lists[b][2].append(ll.numpy())
    ob, os_, ol, oc = fusion.fuse()
    ob, os_, ol, oc = ob.cpu().numpy(), os_.cpu().numpy(), ol.cpu().numpy(), oc.cpu().numpy()
    for b in range(B):
        rb, rs, rl = oracle.wbf.weighted_boxes_fusion(*ref_lists[b], None, 0.55, 0.001)
        m = int(oc[b])
        assert m == len(rs) and m > 0, f"image {b}: fused count {m} vs {len(rs)}"
        assert np.array_equal(ol[b, :m].astype(np.float64), rl)
        assert np.allclose(ob[b, :m].astype(np.float64), rb, rtol=1e-5, atol=1e-7)
        assert np.allclose(os_[b, :m], rs, rtol=1e-5, atol=0)

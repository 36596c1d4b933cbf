"""Size-independent properties at the FULL BASELINE.json sizes (the oracle is too slow there, so correctness is
shown through invariants the domain offers) plus spot checks of a few images against the oracle.

cfg2  YOLOv5s 640^2 B=256 nc=80        sortedness, max_det cap, dense==sparse, idempotence, no surviving overlap,
                                       batch-permutation equivariance, oracle spot check
cfg4  YOLOv5l 1280^2 nc=10 dense       same on the large-image (radix sort + pruned NMS) path
cfg3  FasterRCNN 832^2, 2000 proposals  proposal invariants; RoIAlign linearity / constant-map identity; RoIPool bound
cfg5  TTA + WBF                        single-view identity, duplicate-view invariance
"""
import numpy as np
import pytest
import torch
import torchvision

from _tol import close, boxes_close

pytestmark = pytest.mark.gpu


def _check_detections(det, cnt, idx, max_det, iou_thr, offset=7680.0):
    B = det.shape[0]
    assert int(cnt.max()) <= max_det and int(cnt.min()) >= 0
    ar = torch.arange(max_det, device=det.device)[None, :]
    valid = ar < cnt[:, None]
    sc = det[..., 4]
    d = sc[:, :-1] - sc[:, 1:]
    assert bool((d[valid[:, 1:]] >= 0).all()), "scores must be sorted descending"
    assert bool((idx[valid] >= 0).all())
    # kept boxes of one image never overlap above the threshold on the class-offset boxes (checked on a sample)
    for b in range(0, B, max(B // 8, 1)):
        n = int(cnt[b])
        if n < 2:
            continue
        bx = (det[b, :n, :4] + det[b, :n, 5:6] * offset).cpu()
        iou = torchvision.ops.box_iou(bx, bx)
        iou.fill_diagonal_(0)
        assert float(iou.max()) <= iou_thr + 1e-6


@pytest.fixture(scope="module")
def cfg2_full():
    from heltondetection_b200 import synth
    heads, _ = synth.yolo_heads(256, 640, 80, 20, 1235)
    return heads


def test_cfg2_full_size_invariants(cfg2_full):
    import oracle
    from heltondetection_b200 import yolo, ops
    heads = [h.cuda() for h in cfg2_full]
    pp = yolo.YoloPostprocessor(conf_thres=0.25, iou_thres=0.45, dense_read=True)
    det, cnt, idx = [t.clone() for t in pp(heads)]
    _check_detections(det, cnt, idx, 300, 0.45)
    assert int(cnt.sum()) > 256 * 10
    # objectness-skip mode reads far fewer bytes but must give the same bits
    det2, cnt2, idx2 = yolo.YoloPostprocessor(conf_thres=0.25, iou_thres=0.45, dense_read=False)(heads)
    assert torch.equal(cnt, cnt2)
    m = torch.arange(300, device=det.device)[None, :] < cnt[:, None]
    assert torch.equal(det[m], det2[m]) and torch.equal(idx[m], idx2[m])
    # idempotence: NMS of the kept boxes keeps them all (class-aware, exact class masking)
    for b in (0, 100, 255):
        n = int(cnt[b])
        k = ops.batched_nms(det[b, :n, :4].contiguous(), det[b, :n, 4].contiguous(), det[b, :n, 5].int(), 0.45)
        assert torch.equal(k.cpu(), torch.arange(n))
    # batch-permutation equivariance
    perm = torch.randperm(256, generator=torch.Generator().manual_seed(0)).cuda()
    det3, cnt3, idx3 = yolo.YoloPostprocessor(conf_thres=0.25, iou_thres=0.45)([h[perm].contiguous() for h in heads])
    assert torch.equal(cnt3, cnt[perm])
    assert torch.equal(det3[m[perm]], det[perm][m[perm]])
    # oracle spot check on three images of the full batch
    for b in (3, 128, 250):
        pred = oracle.yolo.decode_box([h[b:b + 1] for h in cfg2_full])
        ref, ridx = oracle.yolo.non_max_suppression(pred, 0.25, 0.45, return_index=True)
        n = int(cnt[b])
        assert torch.equal(idx[b, :n].cpu(), ridx[0]) and boxes_close(det[b, :n, :4], ref[0][:, :4])


def test_cfg4_dense_full_resolution_invariants():
    import oracle
    from heltondetection_b200 import synth, yolo
    heads, _ = synth.yolo_heads(8, 1280, 10, 300, 1238, dense=True)
    pp = yolo.YoloPostprocessor(conf_thres=0.001, iou_thres=0.6)
    det, cnt, idx = pp([h.cuda() for h in heads])
    _check_detections(det, cnt, idx, 300, 0.6)
    assert int(cnt.min()) == 300          # dense scenes fill max_det
    pred = oracle.yolo.decode_box([h[5:6] for h in heads])
    ref, ridx = oracle.yolo.non_max_suppression(pred, 0.001, 0.6, return_index=True)
    assert torch.equal(idx[5, :300].cpu(), ridx[0])


def test_cfg3_full_resolution_invariants():
    from heltondetection_b200 import synth, rpn, ops
    B, img = 2, 832
    obj, dlt, bases, _ = synth.rpn_heads(B, img, G=20, seed=1237)
    pr = rpn.RpnProposals(bases, (4, 8, 16, 32), (img, img), n_pre_nms=12000, n_post_nms=2000, min_size=16)
    rois, cnt, sc, idx = pr([o.cuda() for o in obj], [d.cuda() for d in dlt])
    rois = rois.view(B, 2000, 5)
    assert cnt.tolist() == [2000, 2000]
    assert bool((sc[:, :-1] >= sc[:, 1:]).all())                      # score order
    w, h = rois[..., 3] - rois[..., 1], rois[..., 4] - rois[..., 2]
    assert float(w.min()) >= 16 and float(h.min()) >= 16             # min-size
    assert float(rois[..., 1:].min()) >= 0 and float(rois[..., 1:].max()) <= img   # clipped
    for b in range(B):
        assert len(set(idx[b].tolist())) == 2000                      # distinct proposals
        iou = torchvision.ops.box_iou(rois[b, :, 1:].cpu(), rois[b, :, 1:].cpu())
        iou.fill_diagonal_(0)
        assert float(iou.max()) <= 0.7 + 1e-6                         # nothing kept overlaps a kept box above thr
    # RoIAlign on the real pyramid: linear in the features, exact on constant maps, bounded by RoIPool's max
    feats = [f.cuda() for f in synth.fpn_features(B, img, 256, 1237)]
    feats2 = [f.flip(1).contiguous() for f in feats]
    scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
    r = rois.view(-1, 5)
    fa, lv = ops.multilevel_roi_align(feats, r, 7, scales, 2, False)
    fb, _ = ops.multilevel_roi_align(feats2, r, 7, scales, 2, False)
    fc, _ = ops.multilevel_roi_align([2.0 * a - 0.5 * b2 for a, b2 in zip(feats, feats2)], r, 7, scales, 2, False)
    assert fa.shape == (4000, 256, 7, 7) and int(lv.min()) >= 0 and int(lv.max()) <= 3
    assert torch.allclose(fc, 2.0 * fa - 0.5 * fb, rtol=1e-4, atol=2e-5)
    ones, _ = ops.multilevel_roi_align([torch.full_like(f, 3.0) for f in feats], r, 7, scales, 2, True)
    assert torch.allclose(ones, torch.full_like(ones, 3.0), rtol=1e-6, atol=1e-6)   # weights of a bin sum to 1 inside the map
    ref_lv = torchvision.ops.poolers.LevelMapper(2, 5)([r[:, 1:].cpu()])
    assert torch.equal(lv.cpu(), ref_lv)
    sub = torch.arange(0, 4000, 97)
    ref = torchvision.ops.roi_align(feats[1].cpu(), r[sub].cpu(), 7, 1 / 8, 2, False)
    got = ops.roi_align(feats[1], r[sub], 7, 1 / 8, 2, False)
    assert close(got, ref)


def test_wbf_identities():
    from heltondetection_b200 import wbf
    rng = np.random.default_rng(0)
    n = 40
    c = rng.uniform(0.1, 0.9, (n, 2)); s = rng.uniform(0.01, 0.03, (n, 2))
    b = np.concatenate((c - s, c + s), 1).astype(np.float32)
    sc = rng.uniform(0.1, 1, n).astype(np.float32)
    lb = rng.integers(0, 5, n).astype(np.float32)
    # one view, (almost surely) no overlaps: every box is its own cluster -> output = input sorted by score
    fb, fs, fl = wbf.weighted_boxes_fusion([b], [sc], [lb], None, 0.55, 0.0)
    o = np.argsort(-sc, kind="stable")
    assert np.array_equal(fs, sc[o].astype(np.float64)) and np.array_equal(fl, lb[o].astype(np.float64))
    assert np.allclose(fb, b[o], atol=1e-7)
    # the same view three times: same boxes, same scores (avg of identical scores * min(3,3)/3)
    fb3, fs3, fl3 = wbf.weighted_boxes_fusion([b, b, b], [sc, sc, sc], [lb, lb, lb], None, 0.55, 0.0)
    assert fb3.shape == fb.shape and np.allclose(fb3, fb, atol=1e-6) and np.allclose(fs3, fs, rtol=1e-6)

"""CPU: pins the oracle (tests never trust it blindly).
 - against the committed fixtures generated from the torchvision CPU ops (tests/golden/make_golden.py);
 - independent restatements (oracle/*_restated) against the torchvision ops themselves;
 - the known-answer edge cases of SURVEY.md A.4 / A.5."""
import os
import numpy as np
import pytest
import torch
import torchvision

import oracle

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))


def T(name):
    return torch.from_numpy(G[name])


def test_golden_was_made_with_this_torchvision():
    assert str(G["torchvision_version"]) == torchvision.__version__


@pytest.mark.parametrize("thr", [0.3, 0.5, 0.7])
def test_nms_golden_and_restatement(thr):
    b, s = T("nms_boxes"), T("nms_scores")
    ref = G[f"nms_keep_{thr}"]
    assert np.array_equal(oracle.boxes.nms(b, s, thr).numpy(), ref)
    assert np.array_equal(oracle.boxes.nms_restated(b.numpy(), s.numpy(), thr), ref)


def test_batched_nms_and_iou_golden():
    b, s, c = T("nms_boxes"), T("nms_scores"), T("nms_cls")
    assert np.array_equal(oracle.boxes.batched_nms(b, s, c, 0.5).numpy(), G["batched_keep_0.5"])
    # the coordinate trick and the per-class loop agree here
    assert np.array_equal(torchvision.ops.batched_nms(b, s, c, 0.5).numpy(), G["batched_keep_0.5"])
    assert np.array_equal(oracle.boxes.box_iou(b[:40], b[40:90]).numpy(), G["box_iou_40x50"])
    assert np.allclose(oracle.boxes.box_iou_restated(b[:40].numpy(), b[40:90].numpy()), G["box_iou_40x50"], rtol=1e-6, atol=1e-7)


def test_nms_restated_random_vs_torchvision():
    g = torch.Generator().manual_seed(0)
    for n in (1, 7, 64, 333):
        c = torch.rand((n, 2), generator=g) * 100
        s = torch.rand((n, 2), generator=g) * 40 + 1
        b = torch.cat((c - s / 2, c + s / 2), 1)
        sc = (torch.rand((n,), generator=g) * 8).round() / 8  # ties
        for thr in (0.2, 0.5, 0.6):
            assert np.array_equal(oracle.boxes.nms_restated(b.numpy(), sc.numpy(), thr), torchvision.ops.nms(b, sc, thr).numpy())


def test_nms_known_answers_A4():
    nms = lambda b, s, t: torchvision.ops.nms(torch.tensor(b, dtype=torch.float32), torch.tensor(s, dtype=torch.float32), t).tolist()
    same = [[0, 0, 10, 10]] * 4
    assert nms(same, [1, 1, 1, 1], 0.5) == [0]                       # identical boxes + equal scores -> index 0 only
    assert nms([[5, 5, 5, 5]] * 3, [1, 1, 1], 0.5) == [0, 1, 2]      # zero-area duplicates: 0/0 = NaN -> all kept
    assert nms([[0, 0, 2, 1], [1, 0, 3, 1]], [0.9, 0.8], 1 / 3) == [0]  # IoU = 1/3 as fp32 > double 1/3? pinned below
    b = [[0, 0, 2, 2], [0, 0, 2, 1]]                                  # IoU exactly 0.5
    assert nms(b, [0.9, 0.8], 0.5) == [0, 1]                          # strict >: kept
    assert nms(b, [0.9, 0.8], 0.49) == [0]
    assert nms(same[:2] + [[20, 20, 30, 30]], [0.1, float("nan"), 0.5], 0.5)[0] == 1   # NaN score sorts first
    assert nms([[10, 10, 0, 0], [0, 0, 10, 10]], [0.9, 0.8], 0.1) == [0, 1]            # inverted box: not suppressed
    assert torchvision.ops.nms(torch.zeros((0, 4)), torch.zeros((0,)), 0.5).shape == (0,)
    for args in (same, [1, 1, 1, 1]), ([[0, 0, 2, 1], [1, 0, 3, 1]], [0.9, 0.8]):
        for t in (1 / 3, 0.5, 0.6):
            assert oracle.boxes.nms_restated(np.array(args[0], np.float32), np.array(args[1], np.float32), t).tolist() == nms(args[0], args[1], t)


@pytest.mark.parametrize("sr", [2, 0])
@pytest.mark.parametrize("al", [False, True])
def test_roi_align_golden_and_restatement(sr, al):
    x, rois = T("roi_x"), T("roi_rois")
    ref = G[f"roi_align_sr{sr}_al{int(al)}"]
    assert np.array_equal(oracle.roi.roi_align(x, rois, (7, 7), 0.125, sr, al).numpy(), ref)
    got = oracle.roi.roi_align_restated(x.numpy(), rois.numpy(), (7, 7), 0.125, sr, al)  # incl. out-of-map RoIs
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-5)


def test_roi_pool_golden_and_restatement_bit_exact():
    x, rois = T("roi_x"), T("roi_rois")
    assert np.array_equal(oracle.roi.roi_pool(x, rois, (7, 7), 0.125).numpy(), G["roi_pool"])
    assert np.array_equal(oracle.roi.roi_pool_restated(x.numpy(), rois.numpy(), (7, 7), 0.125), G["roi_pool"])


def test_level_map_known_answer():
    """sides 5 / 111.9 / 112 / 224 / 448 / 900 px: k = floor(4 + log2(s/224) + 1e-6) clamped to [2,5], minus 2."""
    b = T("level_boxes")
    want = [0, 0, 1, 2, 3, 3]
    assert oracle.roi.level_map(b).tolist() == want == G["level_ids"].tolist()
    assert torchvision.ops.poolers.LevelMapper(2, 5)([b]).tolist() == want          # the executable reference
    assert oracle.roi.level_map(b, style="mmdet").tolist() == want                 # mmdet form agrees on the probe
    # exact powers of two sit on the boundary and must round up thanks to eps
    pw = torch.tensor([[0.0, 0.0, 112.0, 112.0], [0.0, 0.0, 111.99, 111.99]])
    assert oracle.roi.level_map(pw).tolist() == [1, 0]


def test_multilevel_oracle_equals_torchvision_pooler():
    from heltondetection_b200 import synth
    feats = synth.fpn_features(2, 128, 8, seed=1)
    rois = synth.random_rois(2, 40, 128, 2)
    out, lv = oracle.roi.multilevel_roi_align(feats, rois, 7, [1 / 4, 1 / 8, 1 / 16, 1 / 32], 2, False)
    pool = torchvision.ops.MultiScaleRoIAlign(["0", "1", "2", "3"], 7, 2)
    boxes = [rois[rois[:, 0] == b, 1:] for b in range(2)]
    ref = pool({str(i): f for i, f in enumerate(feats)}, boxes, [(128, 128)] * 2)
    assert torch.allclose(out, ref, rtol=1e-6, atol=1e-6)


def test_yolo_golden():
    heads = [T(f"yolo_head{l}") for l in range(3)]
    pred = oracle.yolo.decode_box(heads)
    assert np.allclose(pred.numpy(), G["yolo_pred"], rtol=1e-6, atol=1e-6)
    det, idx = oracle.yolo.non_max_suppression(T("yolo_pred"), 0.25, 0.45, return_index=True)
    assert np.array_equal(idx[0].numpy(), G["yolo_idx"]) and np.array_equal(det[0].numpy(), G["yolo_det"])
    assert G["yolo_det"].shape[0] > 0


def test_yolo_decode_formula_spot_check():
    """A.1 by hand on one cell."""
    x = torch.zeros((1, 3 * 6, 2, 2))
    x[0, 6 + 0, 1, 0] = 1.0   # anchor 1, tx at (i=1, j=0)
    x[0, 6 + 2, 1, 0] = -1.0  # tw
    p = oracle.yolo.decode_box([x], (((10, 13), (16, 30), (33, 23)),), (8,))
    row = p[0, (1 * 2 + 1) * 2 + 0]
    sig = lambda v: 1 / (1 + np.exp(-v))
    assert abs(row[0].item() - (2 * sig(1.0) - 0.5 + 0) * 8) < 1e-5 and abs(row[1].item() - (2 * 0.5 - 0.5 + 1) * 8) < 1e-5
    assert abs(row[2].item() - (2 * sig(-1.0)) ** 2 * 16) < 1e-5 and abs(row[3].item() - 30.0) < 1e-5


def test_rpn_golden():
    obj = [T(f"rpn_obj{l}") for l in range(4)]
    dlt = [T(f"rpn_dlt{l}") for l in range(4)]
    bases = [G[f"rpn_base{l}"] for l in range(4)]
    roi, sc, ix = oracle.rpn.rpn_proposals(obj, dlt, bases, (4, 8, 16, 32), (128, 128), n_pre_nms=600, n_post_nms=100, min_size=8)[0]
    assert np.array_equal(ix.numpy(), G["rpn_idx"]) and np.allclose(roi.numpy(), G["rpn_roi"], rtol=1e-6, atol=1e-5)
    assert (roi[:, 2] - roi[:, 0]).min() >= 8 and roi.min() >= 0 and roi.max() <= 128


def test_rpn_anchor_enumeration_matches_torchvision_generator():
    """anchor[(i*W+j)*A+a] = base[a] + shift -- same enumeration as torchvision AnchorGenerator.grid_anchors."""
    base = oracle.rpn.generate_anchor_base(4, (0.5, 1.0, 2.0), (8.0,))
    anc = oracle.rpn.enumerate_shifted_anchor(base, 4, 3, 5)
    assert anc.shape == (3 * 5 * 3, 4)
    assert np.allclose(anc[(2 * 5 + 4) * 3 + 1], base[1] + np.array([16, 8, 16, 8], np.float32))
    assert np.allclose(base[1], [-16, -16, 16, 16])


def test_wbf_golden_and_hand_case():
    bl = [G[f"wbf_b{v}"] for v in range(3)]
    sl = [G[f"wbf_s{v}"] for v in range(3)]
    ll = [G[f"wbf_l{v}"] for v in range(3)]
    b, s, l = oracle.wbf.weighted_boxes_fusion(bl, sl, ll, None, 0.55, 0.1, "avg")
    assert np.array_equal(b, G["wbf_boxes"]) and np.array_equal(s, G["wbf_scores"]) and np.array_equal(l, G["wbf_labels"])
    # two views agree on one box, a third box is alone: fused = score-weighted mean, conf = mean * min(n,V)/V
    b, s, l = oracle.wbf.weighted_boxes_fusion([[[0.1, 0.1, 0.5, 0.5], [0.6, 0.6, 0.9, 0.9]], [[0.1, 0.1, 0.5, 0.5]]],
                                               [[0.8, 0.4], [0.4]], [[0, 1], [0]], None, 0.55, 0.0)
    assert l.tolist() == [0.0, 1.0] and np.allclose(b[0], [0.1, 0.1, 0.5, 0.5], atol=1e-6)
    assert np.allclose(s, [np.float32(0.6), 0.4 * 1 / 2], rtol=1e-6)


def test_tta_map_back_inverts_the_view_transform():
    det = torch.tensor([[100.0, 50.0, 200.0, 150.0, 0.9, 3.0]])
    r, W = 1.2, 768.0
    view = det.clone()
    view[:, [0, 2]] = W - det[:, [2, 0]] * r
    view[:, [1, 3]] = det[:, [1, 3]] * r
    b, s, l = oracle.tta.map_back(view, r, True, W, 640.0, 640.0)
    assert torch.allclose(b * 640, det[:, :4], atol=1e-3) and s.item() == pytest.approx(0.9) and l.item() == 3.0


def test_plain_c_restatement_is_pinned_to_torchvision():
    """oracle/c/hd_oracle.c (gcc, -ffp-contract=off) against the torchvision CPU ops and the golden fixtures."""
    from oracle import cref
    b, s = T("nms_boxes"), T("nms_scores")
    for thr in (0.3, 0.5, 0.7):
        assert np.array_equal(cref.nms(b.numpy(), s.numpy(), thr), G[f"nms_keep_{thr}"])
    g = torch.Generator().manual_seed(3)
    for n in (1, 50, 700):
        c = torch.rand((n, 2), generator=g) * 200
        w = torch.rand((n, 2), generator=g) * 60 + 1
        bb = torch.cat((c - w / 2, c + w / 2), 1)
        sc = (torch.rand((n,), generator=g) * 16).round() / 16
        if n > 10:
            sc[3] = float("nan"); bb[5, 0] = float("nan"); bb[7] = bb[7][[2, 3, 0, 1]]
        for thr in (0.2, 0.5, 0.6):
            assert np.array_equal(cref.nms(bb.numpy(), sc.numpy(), thr), torchvision.ops.nms(bb, sc, thr).numpy())
        assert np.array_equal(cref.nms(bb.numpy(), sc.numpy(), 0.5, max_keep=3), torchvision.ops.nms(bb, sc, 0.5).numpy()[:3])
    assert np.allclose(cref.box_iou(b[:40].numpy(), b[40:90].numpy()), G["box_iou_40x50"], rtol=1e-6, atol=1e-7)
    x, rois = T("roi_x"), T("roi_rois")
    for sr in (2, 0):
        for al in (False, True):
            got = cref.roi_align(x.numpy(), rois.numpy(), 7, 0.125, sr, al)
            assert np.allclose(got, G[f"roi_align_sr{sr}_al{int(al)}"], rtol=1e-6, atol=1e-6)
    assert np.array_equal(cref.roi_pool(x.numpy(), rois.numpy(), 7, 0.125), G["roi_pool"])


def test_roi_head_restatement_equals_torchvision_method():
    """oracle.roi_head.postprocess_detections (restatement with variant switches) == the unmodified torchvision
    RoIHeads.postprocess_detections on the default variant (roi_heads.py:668-723)."""
    import oracle
    g = torch.Generator().manual_seed(0)
    R, C = 300, 21
    lg = torch.randn(2 * R, C, generator=g) * 3
    rg = torch.randn(2 * R, C * 4, generator=g) * 0.5
    xy = torch.rand(2 * R, 2, generator=g) * 300
    pr = torch.cat((xy, xy + torch.rand(2 * R, 2, generator=g) * 150 + 4), 1)
    props, shapes = [pr[:R], pr[R:]], [(400, 420), (380, 450)]
    a = oracle.roi_head.postprocess_detections_tv(lg, rg, props, shapes)
    b = oracle.roi_head.postprocess_detections(lg, rg, props, shapes)
    for x, y in zip(a, b):
        for u, v in zip(x, y):
            assert torch.equal(u, v)
    # letterbox inverse known answer: 640x640 letterbox of a 480x640 image has gain 1 and 80 px of vertical padding
    box = torch.tensor([[10.0, 90.0, 630.0, 700.0]])
    assert oracle.roi_head.scale_coords((640, 640), box, (480, 640)).tolist() == [[10.0, 10.0, 630.0, 480.0]]


def test_roi_head_oracle_against_committed_fixture():
    import os
    import numpy as np
    import oracle
    G2 = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v2.npz"))
    lg, rg, pr = (torch.from_numpy(G2[k]) for k in ("rh_logits", "rh_deltas", "rh_props"))
    b, s, l = oracle.roi_head.postprocess_detections(lg, rg, [pr[:120], pr[120:]], [(256, 320)] * 2, 0.05, 0.5, 50)
    for i in range(2):
        assert np.array_equal(b[i].numpy(), G2[f"rh_boxes{i}"]) and np.array_equal(l[i].numpy(), G2[f"rh_labels{i}"])
        assert np.array_equal(oracle.roi_head.scale_coords((256, 320), b[i], (480, 640)).numpy(), G2[f"rh_scaled{i}"])


def test_plain_c_oracle_match_and_scale_pinned_to_torchvision_and_restatement():
    """oracle/c hdo_match == torchvision det_utils.Matcher(box_iou); hdo_scale_coords == oracle.roi_head.scale_coords"""
    import numpy as np
    import oracle
    from oracle import cref
    from torchvision.models.detection._utils import Matcher
    g = torch.Generator().manual_seed(9)
    def boxes(n):
        xy = torch.rand(n, 2, generator=g) * 500
        return torch.cat((xy, xy + torch.rand(n, 2, generator=g) * 200 + 4), 1)
    gt, pr = boxes(23), boxes(1500)
    pr[:10] = gt[:10]
    for allow in (False, True):
        ref = Matcher(0.7, 0.3, allow)(torchvision.ops.box_iou(gt, pr))
        assert np.array_equal(cref.match(gt.numpy(), pr.numpy(), 0.7, 0.3, allow), ref.numpy())
    det = torch.cat((boxes(64), torch.rand(64, 2, generator=g)), 1)
    gain = min(640 / 1080, 640 / 1920)
    pad = ((640 - 1920 * gain) / 2, (640 - 1080 * gain) / 2)
    ref = oracle.roi_head.scale_coords((640, 640), det[:, :4], (1080, 1920))
    got = cref.scale_coords(det.numpy(), np.float32(pad[0]), np.float32(pad[1]), np.float32(gain), 1920.0, 1080.0)
    assert np.array_equal(got[:, :4], ref.numpy())


def test_round2_oracle_variants_against_golden_v3():
    """the oracle's round-2 variants reproduce tests/golden/golden_v3.npz (exact-math RPN, multi_label, WBF conf types / rescale)"""
    G3 = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v3.npz"))
    obj = [torch.from_numpy(G[f"rpn_obj{l}"]) for l in range(4)]
    dlt = [torch.from_numpy(G[f"rpn_dlt{l}"]) for l in range(4)]
    bases = [G[f"rpn_base{l}"] for l in range(4)]
    roi, sc, ix = oracle.rpn.rpn_proposals(obj, dlt, bases, (4, 8, 16, 32), (128, 128), exact_math=True, n_pre_nms=600, n_post_nms=100, min_size=8)[0]
    assert np.array_equal(ix.numpy(), G3["rpnx_idx"]) and np.array_equal(roi.numpy(), G3["rpnx_roi"])
    # exact_math changes values by at most an ulp of the transcendental: same proposals as the fp32 chain up to near-ties
    roi32, sc32, ix32 = oracle.rpn.rpn_proposals(obj, dlt, bases, (4, 8, 16, 32), (128, 128), n_pre_nms=600, n_post_nms=100, min_size=8)[0]
    assert len(set(ix.tolist()) & set(ix32.tolist())) >= 0.97 * len(ix32)
    heads = [torch.from_numpy(G[f"yolo_head{l}"]) for l in range(3)]
    pred = oracle.yolo.decode_box(heads)
    det, idx = oracle.yolo.non_max_suppression(pred, 0.01, 0.45, return_index=True, multi_label=True)
    assert np.array_equal(idx[0].numpy(), G3["ml_idx_0.01"])
    nc = pred.shape[2] - 5
    x, ai = oracle.yolo.filter_candidates_multi_label(pred[0], 0.01)
    x1, a1 = oracle.yolo.filter_candidates(pred[0], 0.01)
    assert x.shape[0] >= x1.shape[0] and set((ai // nc).tolist()) == set(a1.tolist())       # same anchors, possibly several classes each
    bl = [G[f"wbf_b{v}"] for v in range(3)]
    sl = [G[f"wbf_s{v}"] for v in range(3)]
    ll = [G[f"wbf_l{v}"] for v in range(3)]
    w = G3["wbf_weights"].tolist()
    for ct in ("box_and_model_avg", "absent_model_aware_avg"):
        b, s, l = oracle.wbf.weighted_boxes_fusion(bl, sl, ll, w, 0.55, 0.1, ct)
        assert np.array_equal(s, G3[f"wbf_{ct}_scores"]) and np.array_equal(b, G3[f"wbf_{ct}_boxes"])
    a = oracle.wbf.weighted_boxes_fusion(bl, sl, ll, w, 0.55, 0.1, "avg", False, "len_weights")[1]
    b = oracle.wbf.weighted_boxes_fusion(bl, sl, ll, w, 0.55, 0.1, "avg", False, "sum_weights")[1]
    assert not np.array_equal(a, b)          # the two ensemble-boxes rules differ for non-unit weights ...
    a1 = oracle.wbf.weighted_boxes_fusion(bl, sl, ll, None, 0.55, 0.1, "avg", False, "len_weights")[1]
    b1 = oracle.wbf.weighted_boxes_fusion(bl, sl, ll, None, 0.55, 0.1, "avg", False, "sum_weights")[1]
    assert np.array_equal(a1, b1)            # ... and agree for unit weights

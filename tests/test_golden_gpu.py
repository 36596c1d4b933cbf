"""CUDA path against the committed golden fixtures (tests/golden/golden_v1.npz, generated from the torchvision
CPU ops + oracle restatements by tests/golden/make_golden.py)."""
import os
import numpy as np
import pytest
import torch
from _tol import close, boxes_close

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))


def C(name):
    return torch.from_numpy(G[name]).cuda()


def test_nms_family():
    from heltondetection_b200 import ops
    b, s, c = C("nms_boxes"), C("nms_scores"), C("nms_cls")
    for thr in (0.3, 0.5, 0.7):
        assert np.array_equal(ops.nms(b, s, thr).cpu().numpy(), G[f"nms_keep_{thr}"])
    assert np.array_equal(ops.batched_nms(b, s, c, 0.5).cpu().numpy(), G["batched_keep_0.5"])
    assert np.allclose(ops.box_iou(b[:40], b[40:90]).cpu().numpy(), G["box_iou_40x50"], rtol=1e-6, atol=1e-7)
    assert np.array_equal(torch.ops.hd_b200.nms(b, s, 0.5).cpu().numpy(), G["nms_keep_0.5"])


@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_roi_family(layout):
    from heltondetection_b200 import ops
    x, rois = C("roi_x"), C("roi_rois")
    for sr in (2, 0):
        for al in (False, True):
            got = ops.roi_align(x, rois, (7, 7), 0.125, sr, al, layout=layout)
            assert close(got, torch.from_numpy(G[f"roi_align_sr{sr}_al{int(al)}"]))
    assert np.array_equal(ops.roi_pool(x, rois, (7, 7), 0.125, layout=layout).cpu().numpy(), G["roi_pool"])
    assert ops.level_map(C("level_boxes")).cpu().tolist() == G["level_ids"].tolist()
    o = torch.ops.hd_b200.roi_align(x, rois, 0.125, 7, 7, 2, False)
    assert close(o, torch.from_numpy(G["roi_align_sr2_al0"]))


def test_yolo():
    from heltondetection_b200 import yolo
    heads = [C(f"yolo_head{l}") for l in range(3)]
    pred = yolo.decode_box(heads)
    ref = torch.from_numpy(G["yolo_pred"])
    assert close(pred[..., :4], ref[..., :4]) and close(pred[..., 4:], ref[..., 4:], scale=1e-3)
    det, idx = yolo.postprocess(heads, 0.25, 0.45, return_index=True)
    assert np.array_equal(idx[0].cpu().numpy(), G["yolo_idx"])
    assert boxes_close(det[0][:, :4], torch.from_numpy(G["yolo_det"][:, :4]))
    det2, idx2 = yolo.non_max_suppression(C("yolo_pred"), 0.25, 0.45, return_index=True)
    assert np.array_equal(idx2[0].cpu().numpy(), G["yolo_idx"]) and np.array_equal(det2[0].cpu().numpy(), G["yolo_det"])


def test_rpn():
    from heltondetection_b200 import rpn
    obj = [C(f"rpn_obj{l}") for l in range(4)]
    dlt = [C(f"rpn_dlt{l}") for l in range(4)]
    bases = [G[f"rpn_base{l}"] for l in range(4)]
    pr = rpn.RpnProposals(bases, (4, 8, 16, 32), (128, 128), n_pre_nms=600, n_post_nms=100, min_size=8)
    rois, cnt, sc, idx = pr(obj, dlt)
    n = int(cnt[0])
    want = G["rpn_idx"]
    got = idx[0, :n].cpu().numpy()
    assert len(set(got) & set(want)) >= 0.98 * len(want)
    if np.array_equal(got, want):
        assert boxes_close(rois[:n, 1:], torch.from_numpy(G["rpn_roi"]))


def test_wbf():
    from heltondetection_b200 import wbf
    bl = [G[f"wbf_b{v}"] for v in range(3)]
    sl = [G[f"wbf_s{v}"] for v in range(3)]
    ll = [G[f"wbf_l{v}"] for v in range(3)]
    b, s, l = wbf.weighted_boxes_fusion(bl, sl, ll, None, 0.55, 0.1, "avg")
    assert np.array_equal(l, G["wbf_labels"]) and np.array_equal(b, G["wbf_boxes"]) and np.array_equal(s, G["wbf_scores"])


def test_roi_head_against_torchvision_method_fixture():
    """golden_v2.npz: outputs of torchvision RoIHeads.postprocess_detections (CPU) + scale_coords"""
    from heltondetection_b200 import roi_head
    G2 = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v2.npz"))
    lg, rg, pr = (torch.from_numpy(G2[k]).cuda() for k in ("rh_logits", "rh_deltas", "rh_props"))
    b, s, l = roi_head.postprocess_detections(lg, rg, [pr[:120], pr[120:]], [(256, 320)] * 2, 0.05, 0.5, 50)
    for i in range(2):
        assert np.array_equal(l[i].cpu().numpy(), G2[f"rh_labels{i}"])
        assert boxes_close(b[i], torch.from_numpy(G2[f"rh_boxes{i}"]))
        assert close(s[i], torch.from_numpy(G2[f"rh_scores{i}"]), scale=1e-3)
        k = b[i].shape[0]
        det = torch.zeros((1, k, 6), device="cuda"); det[0, :, :4] = torch.from_numpy(G2[f"rh_boxes{i}"]).cuda()
        sc = roi_head.scale_coords((256, 320), det, [(480, 640)])
        assert np.array_equal(sc[0, :, :4].cpu().numpy(), G2[f"rh_scaled{i}"])


def test_round2_fixtures_golden_v3():
    """golden_v3.npz (tests/golden/make_golden_v3.py): exact-math RPN proposals (torch.equal), multi_label candidates, WBF conf types"""
    from heltondetection_b200 import rpn, yolo, wbf
    G3 = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v3.npz"))
    obj = [C(f"rpn_obj{l}") for l in range(4)]
    dlt = [C(f"rpn_dlt{l}") for l in range(4)]
    bases = [G[f"rpn_base{l}"] for l in range(4)]
    pr = rpn.RpnProposals(bases, (4, 8, 16, 32), (128, 128), n_pre_nms=600, n_post_nms=100, min_size=8, exact_math=True)
    rois, cnt, sc, idx = pr(obj, dlt)
    n = int(cnt[0])
    assert np.array_equal(idx[0, :n].cpu().numpy(), G3["rpnx_idx"])
    assert np.array_equal(rois[:n, 1:].cpu().numpy(), G3["rpnx_roi"]) and np.array_equal(sc[0, :n].cpu().numpy(), G3["rpnx_score"])
    heads = [C(f"yolo_head{l}") for l in range(3)]
    for conf in (0.25, 0.01):
        det, cnt, idx = yolo.YoloPostprocessor(conf_thres=conf, iou_thres=0.45, multi_label=True)(heads)
        n = int(cnt[0])
        assert np.array_equal(idx[0, :n].cpu().numpy(), G3[f"ml_idx_{conf}"])
        assert boxes_close(det[0, :n, :4], torch.from_numpy(G3[f"ml_det_{conf}"][:, :4]))
    bl = [G[f"wbf_b{v}"] for v in range(3)]
    sl = [G[f"wbf_s{v}"] for v in range(3)]
    ll = [G[f"wbf_l{v}"] for v in range(3)]
    w = G3["wbf_weights"].tolist()
    for ct in ("box_and_model_avg", "absent_model_aware_avg"):
        b, s, l = wbf.weighted_boxes_fusion(bl, sl, ll, w, 0.55, 0.1, ct)
        assert np.array_equal(l, G3[f"wbf_{ct}_labels"]) and np.array_equal(b, G3[f"wbf_{ct}_boxes"]) and np.allclose(s, G3[f"wbf_{ct}_scores"], rtol=1e-12, atol=0)
    for rule in ("len_weights", "sum_weights"):
        b, s, l = wbf.weighted_boxes_fusion(bl, sl, ll, w, 0.55, 0.1, "avg", rescale=rule)
        assert np.array_equal(b, G3[f"wbf_avg_{rule}_boxes"]) and np.array_equal(s, G3[f"wbf_avg_{rule}_scores"])

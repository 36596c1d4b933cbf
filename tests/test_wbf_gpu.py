"""WBF / TTA parity against the numpy restatement of ensemble-boxes (oracle/wbf.py).
Cluster membership (counts, labels, order) bit-exact; fused coordinates and scores are reproduced with the
same float64/float32 conventions, asserted <= 1e-5 relative (north_star) and checked for exact equality too."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _views(seed, V=4, n=60, nlab=5, jitter=0.01):
    rng = np.random.default_rng(seed)
    G = 12
    ctr = rng.uniform(0.15, 0.85, (G, 2)); wh = rng.uniform(0.05, 0.3, (G, 2)); glab = rng.integers(0, nlab, G)
    bl, sl, ll = [], [], []
    for v in range(V):
        m = int(rng.integers(n // 2, n))
        pick = rng.integers(0, G, m)
        c = ctr[pick] + rng.normal(0, jitter, (m, 2)); s = wh[pick] * (1 + rng.normal(0, 0.05, (m, 2)))
        b = np.concatenate((c - s / 2, c + s / 2), 1).astype(np.float32)
        bl.append(b); sl.append(rng.uniform(0.05, 1.0, m).astype(np.float32)); ll.append(glab[pick].astype(np.float32))
    return bl, sl, ll


@pytest.mark.parametrize("conf_type", ["avg", "max"])
@pytest.mark.parametrize("weights", [None, [2, 1, 1, 0.5]])
@pytest.mark.parametrize("seed", [0, 1])
def test_wbf_matches_oracle(conf_type, weights, seed):
    import oracle
    from heltondetection_b200 import wbf
    bl, sl, ll = _views(seed)
    rb, rs, rl = oracle.wbf.weighted_boxes_fusion(bl, sl, ll, weights, 0.55, 0.1, conf_type)
    gb, gs, gl = wbf.weighted_boxes_fusion(bl, sl, ll, weights, 0.55, 0.1, conf_type)
    assert gb.shape == rb.shape and np.array_equal(gl, rl)
    assert np.allclose(gb, rb, rtol=1e-5, atol=0) and np.allclose(gs, rs, rtol=1e-5, atol=0)
    assert np.array_equal(gb, rb) and np.array_equal(gs, rs)      # same dtype conventions -> identical bits


@pytest.mark.parametrize("conf_type", ["box_and_model_avg", "absent_model_aware_avg"])
@pytest.mark.parametrize("weights", [None, [2, 1, 1, 0.5]])
def test_wbf_model_aware_conf_types(conf_type, weights):
    """ensemble-boxes >= 1.0.5 conf types: the cluster's distinct models (views) enter the confidence rescale"""
    import oracle
    from heltondetection_b200 import wbf
    bl, sl, ll = _views(3)
    rb, rs, rl = oracle.wbf.weighted_boxes_fusion(bl, sl, ll, weights, 0.55, 0.1, conf_type)
    gb, gs, gl = wbf.weighted_boxes_fusion(bl, sl, ll, weights, 0.55, 0.1, conf_type)
    assert gb.shape == rb.shape and np.array_equal(gl, rl) and np.array_equal(gb, rb)
    assert np.allclose(gs, rs, rtol=1e-12, atol=0)


def test_wbf_rescale_rules_differ_only_for_non_unit_weights():
    import oracle
    from heltondetection_b200 import wbf
    bl, sl, ll = _views(4)
    for weights in (None, [2, 1, 1, 0.5]):
        for rule in ("len_weights", "sum_weights"):
            rb, rs, rl = oracle.wbf.weighted_boxes_fusion(bl, sl, ll, weights, 0.55, 0.1, "avg", False, rule)
            gb, gs, gl = wbf.weighted_boxes_fusion(bl, sl, ll, weights, 0.55, 0.1, "avg", False, rescale=rule)
            assert np.array_equal(gl, rl) and np.array_equal(gb, rb) and np.array_equal(gs, rs)
    a = wbf.weighted_boxes_fusion(bl, sl, ll, None, 0.55, 0.1, rescale="len_weights")
    b = wbf.weighted_boxes_fusion(bl, sl, ll, None, 0.55, 0.1, rescale="sum_weights")
    assert np.array_equal(a[1], b[1])


def test_wbf_edge_cases():
    import oracle
    from heltondetection_b200 import wbf
    # inverted corners, out-of-range coords, zero-area, below-threshold, empty view, allows_overflow, ties
    bl = [np.array([[0.5, 0.5, 0.1, 0.2], [-0.2, 0.1, 0.4, 1.3], [0.3, 0.3, 0.3, 0.6], [0.1, 0.1, 0.4, 0.4], [0.1, 0.1, 0.4, 0.4]], np.float32),
          np.zeros((0, 4), np.float32),
          np.array([[0.11, 0.1, 0.41, 0.4], [0.6, 0.6, 0.9, 0.9]], np.float32)]
    sl = [np.array([0.9, 0.8, 0.7, 0.5, 0.5], np.float32), np.zeros((0,), np.float32), np.array([0.5, 0.01], np.float32)]
    ll = [np.array([1, 1, 0, 2, 2], np.float32), np.zeros((0,), np.float32), np.array([2, 2], np.float32)]
    for ao in (False, True):
        rb, rs, rl = oracle.wbf.weighted_boxes_fusion(bl, sl, ll, None, 0.55, 0.05, "avg", ao)
        gb, gs, gl = wbf.weighted_boxes_fusion(bl, sl, ll, None, 0.55, 0.05, "avg", ao)
        assert np.array_equal(gl, rl) and np.array_equal(gb, rb) and np.array_equal(gs, rs)
    e = wbf.weighted_boxes_fusion([np.zeros((0, 4))], [np.zeros(0)], [np.zeros(0)])
    assert e[0].shape == (0, 4) and e[1].shape == (0,)


def test_wbf_single_label_many_boxes():
    import oracle
    from heltondetection_b200 import wbf
    bl, sl, ll = _views(5, V=6, n=300, nlab=1, jitter=0.02)
    rb, rs, rl = oracle.wbf.weighted_boxes_fusion(bl, sl, ll, None, 0.55, 0.001)
    gb, gs, gl = wbf.weighted_boxes_fusion(bl, sl, ll, None, 0.55, 0.001)
    assert gb.shape == rb.shape and np.array_equal(gb, rb) and np.array_equal(gs, rs)


def test_tta_pipeline_matches_oracle():
    """cfg 5 in miniature: V views of YOLO heads -> per-view decode+NMS -> map back -> WBF."""
    import oracle
    from heltondetection_b200 import synth, yolo, wbf
    B, img, nc = 2, 640, 80
    views, _ = synth.tta_heads(B, img, nc, G=8, seed=1239)
    vspec = [(r, flip, size) for (_, r, flip, size) in views]
    fusion = wbf.TTAFusion(vspec, (img, img), nc, max_det=300, iou_thr=0.55, skip_box_thr=0.001)
    ref_lists = [([], [], []) for _ in range(B)]
    pp = yolo.YoloPostprocessor(conf_thres=0.25, iou_thres=0.45)
    for v, (heads, r, flip, size) in enumerate(views):
        det, cnt, _ = pp([h.cuda() for h in heads])
        fusion.map_back(v, det, cnt)
        # oracle chain on the GPU's per-view detections (stage-wise: WBF inputs identical on both sides)
        for b in range(B):
            d = det[b, : int(cnt[b])].cpu()
            bb, ss, ll = oracle.tta.map_back(d, r, flip, float(size), float(img), float(img))
            ref_lists[b][0].append(bb.numpy()); ref_lists[b][1].append(ss.numpy()); ref_lists[b][2].append(ll.numpy())
    ob, os_, ol, oc = fusion.fuse()
    for b in range(B):
        rb, rs, rl = oracle.wbf.weighted_boxes_fusion(*ref_lists[b], None, 0.55, 0.001)
        m = int(oc[b])
        assert m == len(rs) and m > 0
        assert np.array_equal(ol[b, :m].cpu().numpy().astype(np.float64), rl)
        assert np.allclose(ob[b, :m].cpu().numpy().astype(np.float64), rb, rtol=1e-5, atol=1e-7)
        assert np.allclose(os_[b, :m].cpu().numpy(), rs, rtol=1e-5, atol=0)


def test_tta_run_on_streams_and_graph_equal_the_sequential_chain():
    """TTAFusion.run (views alternating over two side streams, WBF after the join) and its CUDA-graph form return the bits of the
    view-after-view chain, also when the graph is replayed on new head values."""
    from heltondetection_b200 import synth, yolo, wbf
    B, img, nc = 4, 640, 80
    views, _ = synth.tta_heads(B, img, nc, G=8, seed=1240)
    vspec = [(r, flip, size) for (_, r, flip, size) in views]
    devh = [[h.cuda() for h in heads] for (heads, _, _, _) in views]

    def sequential(hv):
        fusion = wbf.TTAFusion(vspec, (img, img), nc, max_det=300, iou_thr=0.55, skip_box_thr=0.001)
        pp = yolo.YoloPostprocessor(conf_thres=0.25, iou_thres=0.45)
        for v in range(len(views)):
            det, cnt, _ = pp(hv[v])
            fusion.map_back(v, det, cnt)
        return [t.clone() for t in fusion.fuse()]

    ref = sequential(devh)
    fusion = wbf.TTAFusion(vspec, (img, img), nc, max_det=300, iou_thr=0.55, skip_box_thr=0.001)
    pps = [yolo.YoloPostprocessor(conf_thres=0.25, iou_thres=0.45) for _ in views]
    got = fusion.run(pps, devh)
    torch.cuda.synchronize()
    for a, b in zip(got, ref):
        assert torch.equal(a, b)
    replay, out = fusion.graph(pps, devh)
    views2, _ = synth.tta_heads(B, img, nc, G=8, seed=1241)
    for hv, (heads, _, _, _) in zip(devh, views2):
        for dst, src in zip(hv, heads):
            dst.copy_(src)
    replay()
    torch.cuda.synchronize()
    ref2 = sequential(devh)
    assert int(ref2[3].sum()) > 0 and not torch.equal(ref2[0], ref[0])
    assert torch.equal(out[3], ref2[3])                      # fused boxes per image
    for b in range(B):                                       # rows past the count keep whatever an earlier call left there
        m = int(ref2[3][b])
        for a, r in zip(out[:3], ref2[:3]):
            assert torch.equal(a[b, :m], r[b, :m])

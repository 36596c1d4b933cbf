"""Parity of the CUDA YOLO path (through the C ABI) against the CPU oracle.

Tolerances: integer results (candidate sets, class ids, keep indices, counts) bit-exact;
floating point <= 1e-5 relative (north_star), written as |a-b| <= 1e-5*max(|ref|,1) for pixel
coordinates / scores.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from _tol import close, boxes_close


def _cfg(B, img, nc, G, seed, dense=False):
    from heltondetection_b200 import synth
    heads, _ = synth.yolo_heads(B, img, nc, G, seed, dense=dense)
    return heads


@pytest.fixture(scope="module")
def cfg1():
    return _cfg(2, 640, 80, 20, 1234)


def test_decode_box_matches_oracle(cfg1):
    import oracle
    from heltondetection_b200 import yolo
    ref = oracle.yolo.decode_box(cfg1)
    got = yolo.decode_box([h.cuda() for h in cfg1])
    assert got.shape == ref.shape
    assert close(got[..., :4], ref[..., :4], scale=1.0)      # pixel coordinates: 1e-5 px floor
    assert close(got[..., 4:], ref[..., 4:], scale=1e-3)     # probabilities


@pytest.mark.parametrize("thr,dense_read,ge", [(0.25, False, False), (0.25, True, False), (0.001, False, False), (0.25, False, True)])
def test_candidates_match_oracle(cfg1, thr, dense_read, ge):
    import oracle
    from heltondetection_b200 import yolo
    pred = oracle.yolo.decode_box(cfg1)
    pp = yolo.YoloPostprocessor(conf_thres=thr, dense_read=dense_read, ge=ge)
    got = pp.candidates([h.cuda() for h in cfg1])
    for b in range(pred.shape[0]):
        ref, ridx = oracle.yolo.filter_candidates(pred[b], thr, ge)
        c, idx = got[b]
        assert torch.equal(idx.cpu(), ridx), f"candidate set differs: {idx.numel()} vs {ridx.numel()}"
        assert torch.equal(c[:, 5].cpu(), ref[:, 5])
        assert boxes_close(c[:, :4], ref[:, :4])
        assert close(c[:, 4], ref[:, 4], scale=1e-3)


@pytest.mark.parametrize("thr,iou,mode", [(0.25, 0.45, "offset"), (0.001, 0.6, "offset"), (0.25, 0.45, "exact"), (0.001, 0.45, "agnostic")])
def test_fused_postprocess_keeps_match_oracle(cfg1, thr, iou, mode):
    import oracle
    from heltondetection_b200 import yolo
    pred = oracle.yolo.decode_box(cfg1)
    agn = mode == "agnostic"
    ref, ridx = oracle.yolo.non_max_suppression(pred, thr, iou, agnostic=agn, class_mode=mode if not agn else "offset", return_index=True)
    got, gidx = yolo.postprocess([h.cuda() for h in cfg1], thr, iou, agnostic=agn, class_mode=mode if not agn else "offset", return_index=True)
    for b in range(len(ref)):
        assert torch.equal(gidx[b].cpu(), ridx[b]), f"keep indices differ for image {b}"
        assert torch.equal(got[b][:, 5].cpu(), ref[b][:, 5])
        assert boxes_close(got[b][:, :4], ref[b][:, :4])
        assert close(got[b][:, 4], ref[b][:, 4], scale=1e-3)


def test_nms_on_decoded_prediction_matches_oracle(cfg1):
    import oracle
    from heltondetection_b200 import yolo
    pred = oracle.yolo.decode_box(cfg1)
    ref, ridx = oracle.yolo.non_max_suppression(pred, 0.25, 0.45, return_index=True)
    got, gidx = yolo.non_max_suppression(pred.cuda(), 0.25, 0.45, return_index=True)
    for b in range(len(ref)):
        # identical fp32 inputs -> identical products, boxes and keeps
        assert torch.equal(gidx[b].cpu(), ridx[b])
        assert torch.equal(got[b].cpu(), ref[b])


def test_dense_visdrone_like():
    import oracle
    from heltondetection_b200 import yolo
    heads = _cfg(1, 1280, 10, 300, 1238, dense=True)
    pred = oracle.yolo.decode_box(heads)
    ref, ridx = oracle.yolo.non_max_suppression(pred, 0.001, 0.6, return_index=True)
    got, gidx = yolo.postprocess([h.cuda() for h in heads], 0.001, 0.6, return_index=True)
    assert torch.equal(gidx[0].cpu(), ridx[0])
    assert boxes_close(got[0][:, :4], ref[0][:, :4])


def test_odd_spatial_size_scalar_path():
    """H*W not divisible by 4 -> scalar-load kernel variant."""
    import oracle
    from heltondetection_b200 import yolo
    g = torch.Generator().manual_seed(5)
    heads = [torch.randn((2, 3 * 9, 7, 9), generator=g), torch.randn((2, 3 * 9, 5, 3), generator=g)]
    anchors = (((10, 13), (16, 30), (33, 23)), ((30, 61), (62, 45), (59, 119)))
    pred = oracle.yolo.decode_box(heads, anchors, (8, 16))
    got = yolo.decode_box([h.cuda() for h in heads], anchors, (8, 16))
    assert close(got[..., :4], pred[..., :4], scale=1.0) and close(got[..., 4:], pred[..., 4:], scale=1e-3)
    ref, ridx = oracle.yolo.non_max_suppression(pred, 0.1, 0.45, return_index=True)
    got, gidx = yolo.postprocess([h.cuda() for h in heads], 0.1, 0.45, anchors=anchors, strides=(8, 16), return_index=True)
    for b in range(2):
        assert torch.equal(gidx[b].cpu(), ridx[b])


def test_no_candidates_and_empty_batch():
    from heltondetection_b200 import yolo
    heads = [torch.full((1, 255, 8, 8), -12.0).cuda()]
    out = yolo.postprocess(heads, 0.25, 0.45, anchors=(((10, 13), (16, 30), (33, 23)),), strides=(8,))
    assert out[0].shape == (0, 6)


@pytest.mark.parametrize("thr", [0.25, 0.02, 0.001])
def test_one_call_entry_equals_separate_entry_points(cfg1, thr):
    """thr 0.25: every image takes the small-image NMS kernel; 0.001: all go to the radix/grid path; 0.02: mixed."""
    from heltondetection_b200 import yolo
    heads = [h.cuda() for h in cfg1]
    a = yolo.YoloPostprocessor(conf_thres=thr, one_call=True)(heads)
    a = [t.clone() for t in a]
    b = yolo.YoloPostprocessor(conf_thres=thr, one_call=False)(heads)
    assert torch.equal(a[1], b[1])
    for i, n in enumerate(a[1].tolist()):
        assert torch.equal(a[0][i, :n], b[0][i, :n]) and torch.equal(a[2][i, :n], b[2][i, :n])


def test_zero_copy_pinned_host_inputs(cfg1):
    from heltondetection_b200 import yolo
    ref = yolo.YoloPostprocessor(conf_thres=0.25)([h.cuda() for h in cfg1])
    ref = [t.clone() for t in ref]
    got = yolo.YoloPostprocessor(conf_thres=0.25, device="cuda:0")([h.pin_memory() for h in cfg1])
    torch.cuda.synchronize()
    assert torch.equal(ref[1], got[1])
    for i, n in enumerate(ref[1].tolist()):
        assert torch.equal(ref[0][i, :n], got[0][i, :n])
    with pytest.raises(RuntimeError, match="CUDA-only"):
        yolo.YoloPostprocessor()(cfg1)          # pageable host memory is refused


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("dense", [False, True])
def test_half_precision_heads_equal_the_fp32_path_on_widened_inputs(dtype, dense):
    """8f-4: 16-bit heads are widened to fp32 on load (exact), so the result must be bit-identical to the fp32
    kernels run on heads.to(dtype).float() -- and through that equal to the oracle on the same widened inputs."""
    import oracle
    from heltondetection_b200 import synth, yolo
    heads, _ = synth.yolo_heads(3, 320, 20, 12, 99)
    h16 = [h.to(dtype).cuda() for h in heads]
    h32 = [h.float() for h in h16]
    pp = yolo.YoloPostprocessor(conf_thres=0.25, iou_thres=0.45, dense_read=dense)
    d16, c16, i16 = [t.clone() for t in pp(h16)]
    d32, c32, i32 = [t.clone() for t in pp(h32)]
    assert torch.equal(c16, c32)
    for b in range(3):
        n = int(c16[b])
        assert torch.equal(d16[b, :n], d32[b, :n]) and torch.equal(i16[b, :n], i32[b, :n])
    ref, ridx = oracle.yolo.non_max_suppression(oracle.yolo.decode_box([h.cpu() for h in h32]), 0.25, 0.45, return_index=True)
    for b in range(3):
        n = int(c16[b])
        assert torch.equal(i16[b, :n].cpu(), ridx[b])
    # odd spatial size -> scalar (non-vectorised) loads
    odd = [torch.randn(2, 3 * 8, 13, 11).to(dtype).cuda(), torch.randn(2, 3 * 8, 7, 5).to(dtype).cuda(), torch.randn(2, 3 * 8, 3, 3).to(dtype).cuda()]
    a = yolo.YoloPostprocessor(conf_thres=0.3, dense_read=dense).candidates(odd)
    bq = yolo.YoloPostprocessor(conf_thres=0.3, dense_read=dense).candidates([o.float() for o in odd])
    for (ca, ia), (cb, ib) in zip(a, bq):
        assert torch.equal(ca, cb) and torch.equal(ia, ib)


@pytest.mark.parametrize("dense", [False, True])
@pytest.mark.parametrize("nc,img", [(80, 320), (10, 416), (1, 160)])
def test_channels_last_heads_give_the_same_detections(dense, nc, img):
    """8f-4: heads in torch.channels_last memory go through the NHWC kernel (same per-(cell,anchor) arithmetic): identical
    candidates, hence bit-identical detections, and equal to the oracle's keep indices."""
    import oracle
    from heltondetection_b200 import synth, yolo
    heads, _ = synth.yolo_heads(3, img, nc, 10, 77 + nc)
    nchw = [h.cuda() for h in heads]
    nhwc = [h.contiguous(memory_format=torch.channels_last) for h in nchw]
    assert all(yolo._is_nhwc(h) for h in nhwc)
    a = yolo.YoloPostprocessor(conf_thres=0.25, iou_thres=0.45, dense_read=dense).candidates(nchw)
    b = yolo.YoloPostprocessor(conf_thres=0.25, iou_thres=0.45, dense_read=dense).candidates(nhwc)
    for (ca, ia), (cb, ib) in zip(a, b):
        assert torch.equal(ia, ib) and torch.equal(ca, cb)
    d1, c1, i1 = [t.clone() for t in yolo.YoloPostprocessor(conf_thres=0.25, iou_thres=0.45, dense_read=dense)(nchw)]
    d2, c2, i2 = [t.clone() for t in yolo.YoloPostprocessor(conf_thres=0.25, iou_thres=0.45, dense_read=dense)(nhwc)]
    assert torch.equal(c1, c2)
    ref, ridx = oracle.yolo.non_max_suppression(oracle.yolo.decode_box(heads), 0.25, 0.45, return_index=True)
    for k in range(3):
        n = int(c1[k])
        assert torch.equal(d1[k, :n], d2[k, :n]) and torch.equal(i1[k, :n], i2[k, :n]) and torch.equal(i2[k, :n].cpu(), ridx[k])


def test_bubbliiiing_per_class_output_order():
    """lineage variant of A.2: conf >= thr, nms per class, detections concatenated class by class"""
    import oracle
    from heltondetection_b200 import synth, yolo
    heads, _ = synth.yolo_heads(2, 320, 20, 14, 55)
    ref = oracle.yolo.non_max_suppression_per_class(oracle.yolo.decode_box(heads), 0.5, 0.4)
    got = yolo.postprocess([h.cuda() for h in heads], 0.5, 0.4, group_by_class=True, class_mode="exact", ge=True, max_det=4096)
    for b in range(2):
        assert got[b].shape == ref[b].shape
        assert torch.equal(got[b][:, 5].cpu(), ref[b][:, 5])
        assert torch.allclose(got[b].cpu(), ref[b], rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("cfg", [(2, 640, 80, 20, 0.001, 0.6, False), (2, 1280, 10, 300, 0.001, 0.6, True), (2, 640, 80, 20, 0.25, 0.45, False)])
def test_multi_label_matches_oracle(cfg):
    """ultralytics multi_label=True (its evaluation setting): every (anchor, class) pair over the threshold is a candidate;
    kept (anchor*nc + class) indices bit-exact, boxes/scores within 1e-5"""
    import oracle
    from heltondetection_b200 import synth, yolo
    from _tol import close, boxes_close
    B, img, nc, G, conf, iou, dense = cfg
    heads, _ = synth.yolo_heads(B, img, nc, G, 77, dense=dense)
    pred = oracle.yolo.decode_box(heads)
    ref, ridx = oracle.yolo.non_max_suppression(pred, conf, iou, return_index=True, multi_label=True)
    ref1 = oracle.yolo.non_max_suppression(pred, conf, iou)
    for dense_read in (True, False):
        pp = yolo.YoloPostprocessor(conf_thres=conf, iou_thres=iou, multi_label=True, dense_read=dense_read)
        det, cnt, idx = pp([h.cuda() for h in heads])
        assert not pp.overflowed()
        for b in range(B):
            n = int(cnt[b])
            assert n == ridx[b].numel() and torch.equal(idx[b, :n].cpu(), ridx[b]), f"image {b}"
            assert boxes_close(det[b, :n, :4], ref[b][:, :4]) and close(det[b, :n, 4], ref[b][:, 4], scale=1e-3)
            assert torch.equal(det[b, :n, 5].cpu(), ref[b][:, 5])
    if conf < 0.01:
        assert any(r.shape[0] != s.shape[0] or not torch.equal(r, s) for r, s in zip(ref, ref1)), "multi_label should change the low-threshold result"


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("dense", [False, True])
@pytest.mark.parametrize("nc,thr", [(1, 0.3), (3, 0.25), (10, 0.001), (11, 0.02), (17, 0.001), (35, 0.01), (82, 0.001)])
def test_packed_survivor_variant_candidates_match_oracle(nc, thr, dense, dtype):
    """conf < 0.05 or nc < 16 selects the packed-survivor decode kernel (objectness + box + first nc % U class planes in one load batch,
    last nc % U planes in one batch, logit-domain pre-test, survivors dealt out one per lane); U = 8 planes for fp32 heads and 16 for
    16-bit heads, so the class counts cover empty / merged / separate head and tail batches for both.  The candidate set, classes and
    indices must equal the oracle's on the widened inputs, in both read modes."""
    import oracle
    from heltondetection_b200 import synth, yolo
    heads, _ = synth.yolo_heads(2, 320, nc, 12, 300 + nc)
    hd = [h.to(dtype).cuda() for h in heads]
    wide = [h.float().cpu() for h in hd]
    pred = oracle.yolo.decode_box(wide)
    got = yolo.YoloPostprocessor(conf_thres=thr, dense_read=dense).candidates(hd)
    n_total = 0
    for b in range(2):
        ref, ridx = oracle.yolo.filter_candidates(pred[b], thr, False)
        c, idx = got[b]
        assert torch.equal(idx.cpu(), ridx), f"candidate set differs: {idx.numel()} vs {ridx.numel()}"
        assert torch.equal(c[:, 5].cpu(), ref[:, 5])
        assert boxes_close(c[:, :4], ref[:, :4])
        assert close(c[:, 4], ref[:, 4], scale=1e-3)
        n_total += ridx.numel()
    assert n_total > 0


@pytest.mark.parametrize("cycle_exact,min_cycle", [(False, 4), (True, 5)])
def test_postprocess_pipeline_cycles_equal_step_by_step(cycle_exact, min_cycle):
    """PostprocessPipeline (steps over `depth` streams; whole cycles as one CUDA graph, the remainder step by step): after n steps the
    postprocessor that ran step k last holds the detections of pool[k % len(pool)] -- for the rounded cycle length (a multiple of
    the input/workspace period) and for an exact one that is not a multiple of anything."""
    from heltondetection_b200 import synth, yolo
    pool = [[h.cuda() for h in synth.yolo_heads(2, 320, 20, 8, 500 + j)[0]] for j in range(3)]
    single = yolo.YoloPostprocessor(conf_thres=0.25, iou_thres=0.45)
    want = []
    for hs in pool:
        want.append([t.clone() for t in single(hs)])
    pipe = yolo.PostprocessPipeline(pool, depth=2, cycle_graph=True, min_cycle=min_cycle, cycle_exact=cycle_exact, conf_thres=0.25, iou_thres=0.45)
    assert pipe.cycle is not None, getattr(pipe, "cycle_error", "")
    assert pipe.cycle_len == (5 if cycle_exact else 6)
    for n in (pipe.cycle_len, pipe.cycle_len + 2, 2 * pipe.cycle_len + 1):
        for t in pipe.outputs(0) + pipe.outputs(1):
            t.zero_()
        pipe.fork(); pipe.run(0, n); pipe.join()
        torch.cuda.synchronize()
        # which step ran last on each of the two postprocessors: cycles always execute steps 0..cycle_len-1, the remainder continues at k
        last = {}
        k = 0
        while n - k >= pipe.cycle_len:
            for kk in range(pipe.cycle_len):
                last[kk % 2] = kk % 3
            k += pipe.cycle_len
        while k < n:
            g = k % pipe.n_graphs
            last[g % 2] = g % 3
            k += 1
        for j in (0, 1):
            det, cnt, idx = pipe.outputs(j)
            wd, wc, wi = want[last[j]]
            assert torch.equal(cnt, wc)
            for b in range(2):
                m = int(wc[b])
                assert m > 0 and torch.equal(det[b, :m], wd[b, :m]) and torch.equal(idx[b, :m], wi[b, :m])

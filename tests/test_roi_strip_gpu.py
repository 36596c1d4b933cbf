"""Row-walk (default) and streamed (TMA ring, opt-in) RoIAlign kernels against the per-RoI gather kernels and torchvision's CPU op."""
import pytest
import torch
import torchvision

pytestmark = pytest.mark.gpu


def _rois(B, per_img, img, seed, big=0.15):
    g = torch.Generator().manual_seed(seed)
    n = B * per_img
    side = torch.exp(torch.rand(n, generator=g) * 3.2 + 2.3)            # ~10 .. 245 px
    side = torch.where(torch.rand(n, generator=g) < big, side * 3, side)
    w = side * torch.exp((torch.rand(n, generator=g) - 0.5) * 1.2)
    h = side * torch.exp((torch.rand(n, generator=g) - 0.5) * 1.2)
    cx, cy = torch.rand(n, generator=g) * img, torch.rand(n, generator=g) * img
    r = torch.stack((torch.arange(B).repeat_interleave(per_img).float(), cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2), 1)
    r[::97, 1:] += torch.tensor([-60.0, -60.0, 40.0, 40.0])            # some reach outside the image
    r[::131, 3] = r[::131, 1] + 0.5                                    # degenerate (thinner than a bin)
    return r


@pytest.mark.parametrize("mode", [3, 2])
@pytest.mark.parametrize("aligned", [False, True])
@pytest.mark.parametrize("sr", [2, 1])
def test_strip_equals_gather_multilevel(sr, aligned, mode):
    from heltondetection_b200 import ops, roi
    B, img, C = 3, 416, 64
    g = torch.Generator().manual_seed(5)
    feats = [torch.randn((B, C, img // s, img // s), generator=g).cuda().contiguous(memory_format=torch.channels_last) for s in (4, 8, 16, 32)]
    rois = _rois(B, 700, img, 11).cuda()
    scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
    old = roi.set_mode(mode)       # 3: row-walk kernel over the bucketed RoIs, 2: streamed kernel
    try:
        a, la = ops.multilevel_roi_align(feats, rois, 7, scales, sr, aligned)
    finally:
        roi.set_mode(old)
    old = roi.set_mode(1)
    try:
        b, lb = ops.multilevel_roi_align(feats, rois, 7, scales, sr, aligned)
    finally:
        roi.set_mode(old)
    assert torch.equal(la, lb)
    assert torch.equal(a, b), f"max abs diff {(a - b).abs().max().item():.3e}, rows differing {(a != b).flatten(1).any(1).sum().item()}"


@pytest.mark.parametrize("mode", [3, 2])
@pytest.mark.parametrize("PH,PW", [(7, 7), (5, 3), (6, 6)])
def test_strip_single_level_vs_torchvision_cpu(PH, PW, mode):
    from heltondetection_b200 import ops, roi
    B, H, W, C = 2, 120, 300, 32            # W > one strip: exercises the strip halo
    g = torch.Generator().manual_seed(9)
    x = torch.randn((B, C, H, W), generator=g)
    rois = _rois(B, 600, 1200, 3)
    rois[:, 2] *= 0.4; rois[:, 4] *= 0.4
    old = roi.set_mode(mode)
    try:
        got = ops.roi_align(x.cuda().contiguous(memory_format=torch.channels_last), rois.cuda(), (PH, PW), 0.25, 2, False)
    finally:
        roi.set_mode(old)
    ref = torchvision.ops.roi_align(x, rois, (PH, PW), 0.25, 2, False)
    d = (got.cpu().double() - ref.double()).abs()
    assert bool((d <= 1e-5 * ref.double().abs().clamp(min=1.0)).all()), float(d.max())

"""Label assignment (8f-3) against torchvision's own Matcher over box_iou, both on CPU: bit-exact int64 matches."""
import pytest
import torch
import torchvision
from torchvision.models.detection._utils import Matcher as TVMatcher

pytestmark = pytest.mark.gpu


def _boxes(n, seed, img=800.0, lo=8.0, hi=300.0):
    g = torch.Generator().manual_seed(seed)
    xy = torch.rand(n, 2, generator=g) * (img - hi)
    wh = torch.rand(n, 2, generator=g) * (hi - lo) + lo
    return torch.cat((xy, xy + wh), 1)


@pytest.mark.parametrize("G,N", [(1, 50), (7, 5000), (300, 20000), (600, 1000)])
@pytest.mark.parametrize("high,low,allow", [(0.7, 0.3, True), (0.5, 0.5, False), (0.7, 0.3, False)])
def test_matcher_bit_exact(G, N, high, low, allow):
    from heltondetection_b200 import assign
    gt, an = _boxes(G, 1 + G), _boxes(N, 2 + N)
    an[: min(G, N) // 2] = gt[: min(G, N) // 2]          # exact duplicates: IoU 1 and ties between GTs
    if G > 2:
        gt[2] = gt[1]                                    # two identical GTs: torch.max keeps the first
    ref = TVMatcher(high, low, allow)(torchvision.ops.box_iou(gt, an))
    got, iou = assign.Matcher(high, low, allow)(gt.cuda(), an.cuda(), return_iou=True)
    assert torch.equal(got.cpu(), ref)
    assert torch.equal(iou.cpu(), torchvision.ops.box_iou(gt, an).max(0)[0])


def test_matcher_batched_padded_and_labels():
    from heltondetection_b200 import assign
    B, Gmax, N = 3, 40, 3000
    an = _boxes(N, 5)
    gts = [_boxes(g, 10 + g) for g in (40, 13, 1)]
    gt = torch.zeros(B, Gmax, 4)
    for b, g in enumerate(gts):
        gt[b, : g.shape[0]] = g
    cnt = torch.tensor([40, 13, 1], dtype=torch.int32)
    got = assign.Matcher(0.7, 0.3, True)(gt.cuda(), an.cuda(), cnt.cuda())
    props = torch.stack([_boxes(N, 20 + b) for b in range(B)])
    gotp = assign.Matcher(0.5, 0.5, False)(gt.cuda(), props.cuda(), cnt.cuda())
    for b in range(B):
        assert torch.equal(got[b].cpu(), TVMatcher(0.7, 0.3, True)(torchvision.ops.box_iou(gts[b], an)))
        assert torch.equal(gotp[b].cpu(), TVMatcher(0.5, 0.5, False)(torchvision.ops.box_iou(gts[b], props[b])))
    arg, label = assign.anchor_labels(gts[0].cuda(), an.cuda())
    ref = TVMatcher(0.7, 0.3, True)(torchvision.ops.box_iou(gts[0], an))
    assert torch.equal(label.cpu(), torch.where(ref >= 0, 1, torch.where(ref == -1, 0, -1)))
    with pytest.raises(ValueError):
        assign.Matcher(0.7, 0.3)(torch.zeros(0, 4).cuda(), an.cuda())


def test_box_encode_matches_torchvision_boxcoder():
    """regression targets (BoxCoder.encode_single) on the matched pairs: <= 1e-5 relative (logf differs by ulps between libm and CUDA)"""
    from torchvision.models.detection._utils import BoxCoder
    from heltondetection_b200 import assign
    gt, an = _boxes(25, 3), _boxes(4000, 4)
    m = TVMatcher(0.7, 0.3, True)(torchvision.ops.box_iou(gt, an))
    for w in ((1.0, 1.0, 1.0, 1.0), (10.0, 10.0, 5.0, 5.0)):
        ref = BoxCoder(w).encode_single(gt[m.clamp(min=0)], an)
        got = assign.encode_boxes(gt.cuda(), an.cuda(), m.cuda(), w).cpu()
        assert bool(((got.double() - ref.double()).abs() <= 1e-5 * ref.double().abs().clamp(min=1e-2)).all())
    ref = BoxCoder((1.0, 1.0, 1.0, 1.0)).encode_single(gt, an[:25])
    assert bool(((assign.encode_boxes(gt.cuda(), an[:25].cuda()).cpu() - ref).abs() <= 1e-5 * ref.abs().clamp(min=1e-2)).all())

"""CPU property tests (hypothesis) of the oracle's independent restatements against the torchvision CPU ops."""
import numpy as np
import torch
import torchvision
from hypothesis import given, settings, strategies as st

import oracle

coord = st.floats(min_value=-50, max_value=300, allow_nan=False, width=32)
score = st.sampled_from([0.1, 0.25, 0.5, 0.5, 0.75, 0.9]) | st.floats(min_value=0, max_value=1, width=32)


@settings(max_examples=60, deadline=None)
@given(st.lists(st.tuples(coord, coord, coord, coord, score), min_size=0, max_size=40), st.sampled_from([0.0, 0.3, 0.5, 0.6, 0.7]))
def test_nms_restated_equals_torchvision(rows, thr):
    a = np.array(rows, np.float32).reshape(-1, 5)
    b, s = torch.from_numpy(a[:, :4].copy()), torch.from_numpy(a[:, 4].copy())
    ref = torchvision.ops.nms(b, s, thr).numpy()
    assert np.array_equal(oracle.boxes.nms_restated(b.numpy(), s.numpy(), thr), ref)


@settings(max_examples=40, deadline=None)
@given(st.lists(st.tuples(coord, coord, st.floats(min_value=0.5, max_value=120, width=32), st.floats(min_value=0.5, max_value=120, width=32)),
                min_size=1, max_size=6),
       st.sampled_from([1, 2, 0]), st.booleans())
def test_roi_align_restated_equals_torchvision(rows, sr, aligned):
    g = torch.Generator().manual_seed(len(rows))
    x = torch.randn((1, 3, 9, 11), generator=g)
    r = np.array([[0, x1, y1, x1 + w, y1 + h] for x1, y1, w, h in rows], np.float32)
    ref = torchvision.ops.roi_align(x, torch.from_numpy(r), (3, 4), 0.1, sr, aligned).numpy()
    got = oracle.roi.roi_align_restated(x.numpy(), r, (3, 4), 0.1, sr, aligned)
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-5)
    assert np.array_equal(oracle.roi.roi_pool_restated(x.numpy(), r, (3, 4), 0.1), torchvision.ops.roi_pool(x, torch.from_numpy(r), (3, 4), 0.1).numpy())


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 3), st.integers(1, 40), st.integers(2, 12), st.integers(0, 10_000), st.sampled_from([0.05, 0.3, 0.6]), st.sampled_from([0.3, 0.5, 0.7]))
def test_roi_head_restatement_equals_torchvision_method_on_random_shapes(B, R, C, seed, score_thr, nms_thr):
    """8f-1 oracle: the restatement (with its variant switches at their defaults) == torchvision's RoIHeads.postprocess_detections"""
    import torch
    import oracle
    g = torch.Generator().manual_seed(seed)
    lg = torch.randn(B * R, C, generator=g) * 3
    rg = torch.randn(B * R, C * 4, generator=g) * 0.6
    xy = torch.rand(B * R, 2, generator=g) * 200
    pr = torch.cat((xy, xy + torch.rand(B * R, 2, generator=g) * 100 + 1), 1)
    props, shapes = [pr[b * R:(b + 1) * R] for b in range(B)], [(240, 260)] * B
    a = oracle.roi_head.postprocess_detections_tv(lg, rg, props, shapes, score_thr, nms_thr, 20)
    b = oracle.roi_head.postprocess_detections(lg, rg, props, shapes, score_thr, nms_thr, 20)
    for x, y in zip(a, b):
        for u, v in zip(x, y):
            assert torch.equal(u, v)

"""The FasterRCNN post-CNN chain end to end on the GPU -- RPN proposals -> multi-level RoIAlign -> (a stand-in linear box head) ->
RoI-head post-process -> letterbox inverse -- against the same chain on the CPU (oracle / torchvision ops, same head weights).

The stages are tested exactly one by one elsewhere; here the point is that they compose with no host round trip in between
(padded [B*R] layouts, counts on the device).  The CPU chain starts from the GPU's stage-1 RPN arrays and proposals (so ulp-level
sigmoid differences cannot reshuffle the proposal set) and the features differ by <= 1e-5, so the final detections must agree
except for candidates within ~1e-4 of a threshold: >= 98 % of the detection ids must coincide."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_faster_rcnn_postprocess_chain():
    import oracle
    from heltondetection_b200 import ops, roi_head, rpn, synth
    B, img, C, NCLS = 2, 256, 32, 7
    obj, dlt, bases, _ = synth.rpn_heads(B, img, G=8, seed=1311)
    feats = synth.fpn_features(B, img, C, seed=1312)
    scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
    g = torch.Generator().manual_seed(1313)
    w_cls = torch.randn(C * 49, NCLS, generator=g) * 0.05
    w_box = torch.randn(C * 49, NCLS * 4, generator=g) * 0.01
    # ---- GPU chain: everything stays padded on the device
    pr = rpn.RpnProposals(bases, (4, 8, 16, 32), (img, img), n_pre_nms=2000, n_post_nms=200, min_size=8)
    rois, cnt, sc, idx = pr([o.cuda() for o in obj], [d.cuda() for d in dlt])                     # [B*200, 5]
    f, lv = ops.multilevel_roi_align([x.cuda().contiguous(memory_format=torch.channels_last) for x in feats], rois, 7, scales, 2, False)
    flat = f.flatten(1)
    logits, deltas = flat @ w_cls.cuda(), flat @ w_box.cuda()
    post = roi_head.RoIHeadPostprocessor((img, img), 0.05, 0.5, 50)
    det, did, dcnt = post(logits, deltas, rois, cnt)
    out = roi_head.scale_coords((img, img), det, [(480, 640)] * B, dcnt)
    # ---- CPU chain on the same proposals
    r_cpu, n_cpu = rois.cpu(), cnt.cpu().tolist()
    props = [r_cpu[b * 200: b * 200 + n_cpu[b], 1:] for b in range(B)]
    r5 = torch.cat([r_cpu[b * 200: b * 200 + n_cpu[b]] for b in range(B)])
    fr, _ = oracle.roi.multilevel_roi_align(feats, r5, 7, scales, 2, False)
    fl = fr.flatten(1)
    rb, rs, rl, ri = oracle.roi_head.postprocess_detections(fl @ w_cls, fl @ w_box, props, [(img, img)] * B, 0.05, 0.5, 50, return_ids=True)
    agree = total = 0
    for b in range(B):
        k = int(dcnt[b])
        got, want = set(did[b, :k].cpu().tolist()), set(ri[b].tolist())
        agree += len(got & want); total += max(len(want), 1)
        if torch.equal(did[b, :k].cpu(), ri[b]):
            ref = oracle.roi_head.scale_coords((img, img), rb[b], (480, 640))
            assert torch.allclose(out[b, :k, :4].cpu(), ref, rtol=1e-4, atol=1e-2)
    assert agree / total >= 0.98, (agree, total)

#!/usr/bin/env python
"""Generates tests/golden/golden_v2.npz: the FasterRCNN final stage, produced by torchvision's own
RoIHeads.postprocess_detections (models/detection/roi_heads.py:668-723) on CPU, plus the letterbox inverse
(ultralytics scale_coords).  Re-run:  python tests/golden/make_golden_v2.py"""
import os
import sys
import numpy as np
import torch
import torchvision

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402


def main():
    g = torch.Generator().manual_seed(77)
    B, R, C = 2, 120, 9
    lg = torch.randn(B * R, C, generator=g) * 3
    rg = torch.randn(B * R, C * 4, generator=g) * 0.6
    xy = torch.rand(B * R, 2, generator=g) * 200
    pr = torch.cat((xy, xy + torch.rand(B * R, 2, generator=g) * 120 + 4), 1)
    props, shapes = [pr[:R], pr[R:]], [(256, 320), (256, 320)]
    b, s, l = oracle.roi_head.postprocess_detections_tv(lg, rg, props, shapes, 0.05, 0.5, 50)
    out = {"torchvision_version": np.array(torchvision.__version__), "rh_logits": lg.numpy(), "rh_deltas": rg.numpy(), "rh_props": pr.numpy()}
    for i in range(B):
        out[f"rh_boxes{i}"], out[f"rh_scores{i}"], out[f"rh_labels{i}"] = b[i].numpy(), s[i].numpy(), l[i].numpy()
        out[f"rh_scaled{i}"] = oracle.roi_head.scale_coords((256, 320), b[i], (480, 640)).numpy()
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v2.npz"), **out)
    print("wrote golden_v2.npz", {k: v.shape for k, v in out.items() if k != "torchvision_version"})


if __name__ == "__main__":
    main()

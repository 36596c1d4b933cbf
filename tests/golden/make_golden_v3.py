#!/usr/bin/env python
"""Generates tests/golden/golden_v3.npz: fixtures for the round-2 semantics.

  rpnx_*   RPN proposals of the lineage ProposalCreator chain with exact_math (fp64 transcendentals rounded once) on the inputs of
           golden_v1 (rpn_obj*/rpn_dlt*)
  ml_*     ultralytics multi_label candidates + NMS on the golden_v1 YOLO heads
  wbf_*    ensemble-boxes conf types box_and_model_avg / absent_model_aware_avg and both 'avg' rescale rules (non-unit weights)
Re-run:  python tests/golden/make_golden_v3.py      (the reference's own source is not in the mount: see make_golden.py)
"""
import os
import sys
import numpy as np
import torch
import torchvision

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    G = np.load(os.path.join(here, "golden_v1.npz"))
    out = {"torchvision_version": np.array(torchvision.__version__), "torch_version": np.array(torch.__version__)}
    obj = [torch.from_numpy(G[f"rpn_obj{l}"]) for l in range(4)]
    dlt = [torch.from_numpy(G[f"rpn_dlt{l}"]) for l in range(4)]
    bases = [G[f"rpn_base{l}"] for l in range(4)]
    roi, sc, ix = oracle.rpn.rpn_proposals(obj, dlt, bases, (4, 8, 16, 32), (128, 128), exact_math=True, n_pre_nms=600, n_post_nms=100, min_size=8)[0]
    out["rpnx_roi"], out["rpnx_score"], out["rpnx_idx"] = roi.numpy(), sc.numpy(), ix.numpy()
    heads = [torch.from_numpy(G[f"yolo_head{l}"]) for l in range(3)]
    pred = oracle.yolo.decode_box(heads)
    for conf in (0.25, 0.01):
        det, idx = oracle.yolo.non_max_suppression(pred, conf, 0.45, return_index=True, multi_label=True)
        out[f"ml_det_{conf}"], out[f"ml_idx_{conf}"] = det[0].numpy(), idx[0].numpy()
    bl = [G[f"wbf_b{v}"] for v in range(3)]
    sl = [G[f"wbf_s{v}"] for v in range(3)]
    ll = [G[f"wbf_l{v}"] for v in range(3)]
    w = [2.0, 1.0, 0.5]
    for ct in ("box_and_model_avg", "absent_model_aware_avg"):
        b, s, l = oracle.wbf.weighted_boxes_fusion(bl, sl, ll, w, 0.55, 0.1, ct)
        out[f"wbf_{ct}_boxes"], out[f"wbf_{ct}_scores"], out[f"wbf_{ct}_labels"] = b, s, l
    for rule in ("len_weights", "sum_weights"):
        b, s, l = oracle.wbf.weighted_boxes_fusion(bl, sl, ll, w, 0.55, 0.1, "avg", False, rule)
        out[f"wbf_avg_{rule}_boxes"], out[f"wbf_avg_{rule}_scores"], out[f"wbf_avg_{rule}_labels"] = b, s, l
    out["wbf_weights"] = np.array(w)
    path = os.path.join(here, "golden_v3.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

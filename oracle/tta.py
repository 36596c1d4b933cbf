"""TTA view map-back oracle (README.md:19).

A view is (scale r, hflip).  Detections [n,6] in view pixel coords are un-flipped
(x' = W_view - x, corners swapped), un-scaled (/ r) and normalised to [0,1] by the
original image size -- the input format of weighted_boxes_fusion."""
import torch

DEFAULT_VIEWS = ((0.83, False), (0.83, True), (1.0, False), (1.0, True), (1.17, False), (1.17, True))


def map_back(det, scale, hflip, view_w, img_w, img_h):
    det = det.cpu().float()
    x1, y1, x2, y2 = det[:, 0], det[:, 1], det[:, 2], det[:, 3]
    if hflip:
        x1, x2 = view_w - x2, view_w - x1
    b = torch.stack((x1 / scale / img_w, y1 / scale / img_h, x2 / scale / img_w, y2 / scale / img_h), 1)
    return b, det[:, 4], det[:, 5]

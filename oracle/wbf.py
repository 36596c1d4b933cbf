"""Weighted Boxes Fusion oracle.

Reference feature: README.md:19 (TTA fused by WBF).  The dependency is presumably
ZFTurbo ``ensemble-boxes`` (PyPI; version unknown; NOT installed here) -- PARITY
UNPINNED.  This is a restatement of its published ``weighted_boxes_fusion``
(conf types 'avg', 'max', 'box_and_model_avg', 'absent_model_aware_avg'; the 'avg' rescale of
releases >= 1.0.5, ``min(len(weights), n)``, with ``rescale="sum_weights"`` for the older
``min(weights.sum(), n)``), keeping its dtype conventions: rows are float64,
the fused box is accumulated in a float32 ``np.zeros(8)`` from float64 products,
the confidence sum is float64, IoU matching is float64 with a strict ``>``.
One deliberate fix-point: the original sorts with ``argsort()[::-1]`` (unstable);
here ties are ordered as ``argsort(kind='stable')[::-1]`` (later index first).
"""
import numpy as np


def prefilter_boxes(boxes, scores, labels, weights, thr):
    new_boxes = {}
    for t in range(len(boxes)):
        b_t = np.asarray(boxes[t], np.float64).reshape(-1, 4)
        for j in range(len(b_t)):
            score = float(scores[t][j])
            if score < thr:
                continue
            label = int(labels[t][j])
            x1, y1, x2, y2 = (float(v) for v in b_t[j])
            if x2 < x1:
                x1, x2 = x2, x1
            if y2 < y1:
                y1, y2 = y2, y1
            x1, y1, x2, y2 = (min(max(v, 0.0), 1.0) for v in (x1, y1, x2, y2))
            if (x2 - x1) * (y2 - y1) == 0.0:
                continue
            new_boxes.setdefault(label, []).append(
                [label, score * float(weights[t]), float(weights[t]), t, x1, y1, x2, y2])
    for k in new_boxes:
        cur = np.array(new_boxes[k], np.float64)
        new_boxes[k] = cur[cur[:, 1].argsort(kind="stable")[::-1]]
    return new_boxes


def get_weighted_box(boxes, conf_type="avg"):
    box = np.zeros(8, dtype=np.float32)
    conf = np.float64(0)
    conf_list = []
    w = np.float64(0)
    for b in boxes:
        box[4:] += (b[1] * b[4:])          # f64 product, f64 add, stored f32
        conf += b[1]
        conf_list.append(b[1])
        w += b[2]
    box[0] = boxes[0][0]
    if conf_type == "max":
        box[1] = np.array(conf_list).max()
    else:
        box[1] = conf / len(boxes)
    box[2] = w
    box[3] = -1
    box[4:] /= conf                          # f64 divide, stored f32
    return box


def _find_matching_box(weighted, new_box, match_iou):
    if weighted.shape[0] == 0:
        return -1, match_iou
    b = weighted[:, 4:]
    nb = new_box[4:]
    xA = np.maximum(b[:, 0], nb[0])
    yA = np.maximum(b[:, 1], nb[1])
    xB = np.minimum(b[:, 2], nb[2])
    yB = np.minimum(b[:, 3], nb[3])
    inter = np.maximum(xB - xA, 0) * np.maximum(yB - yA, 0)
    areaA = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    areaB = (nb[2] - nb[0]) * (nb[3] - nb[1])
    with np.errstate(divide="ignore", invalid="ignore"):
        ious = inter / (areaA + areaB - inter)
    ious[weighted[:, 0] != new_box[0]] = -1
    best = int(np.argmax(ious))
    if ious[best] <= match_iou:
        return -1, match_iou
    return best, ious[best]


def weighted_boxes_fusion(boxes_list, scores_list, labels_list, weights=None, iou_thr=0.55,
                          skip_box_thr=0.0, conf_type="avg", allows_overflow=False, rescale="len_weights"):
    """-> (boxes [m,4] f64, scores [m] f64, labels [m] f64), sorted by score desc."""
    if weights is None:
        weights = np.ones(len(boxes_list))
    weights = np.array(weights, np.float64)
    filtered = prefilter_boxes(boxes_list, scores_list, labels_list, weights, skip_box_thr)
    if len(filtered) == 0:
        return np.zeros((0, 4)), np.zeros((0,)), np.zeros((0,))
    overall = []
    for label in filtered:
        boxes = filtered[label]
        clusters = []
        weighted = np.empty((0, 8))
        for j in range(len(boxes)):
            idx, _ = _find_matching_box(weighted, boxes[j], iou_thr)
            if idx != -1:
                clusters[idx].append(boxes[j])
                weighted[idx] = get_weighted_box(clusters[idx], conf_type)
            else:
                clusters.append([boxes[j].copy()])
                weighted = np.vstack((weighted, boxes[j].copy()))
        for i in range(len(clusters)):
            if conf_type == "box_and_model_avg":
                cb = np.array(clusters[i])
                weighted[i, 1] = weighted[i, 1] * len(cb) / weighted[i, 2]
                _, idx = np.unique(cb[:, 3], return_index=True)
                weighted[i, 1] = weighted[i, 1] * cb[idx, 2].sum() / weights.sum()
            elif conf_type == "absent_model_aware_avg":
                cb = np.array(clusters[i])
                models = np.unique(cb[:, 3]).astype(int)
                mask = np.ones(len(weights), dtype=bool)
                mask[models] = False
                weighted[i, 1] = weighted[i, 1] * len(cb) / (weighted[i, 2] + weights[mask].sum())
            elif conf_type == "max":
                weighted[i, 1] = weighted[i, 1] / weights.max()
            elif not allows_overflow:
                cap = weights.sum() if rescale == "sum_weights" else len(weights)
                weighted[i, 1] = weighted[i, 1] * min(len(clusters[i]), cap) / weights.sum()
            else:
                weighted[i, 1] = weighted[i, 1] * len(clusters[i]) / weights.sum()
        overall.append(weighted)
    overall = np.concatenate(overall, axis=0)
    overall = overall[overall[:, 1].argsort(kind="stable")[::-1]]
    return overall[:, 4:], overall[:, 1], overall[:, 0]

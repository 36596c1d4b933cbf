"""CPU: the C-ABI library loads, exports every symbol include/hd_b200.h declares, validates arguments
without touching a GPU, and the host layer refuses CPU tensors (no fallback)."""
import ctypes
import os
import re
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "hd_b200.h")).read()
    return sorted(set(re.findall(r"HD_API\s+[\w\s\*]+?\b(hd_\w+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    from heltondetection_b200 import _lib
    names = _declared()
    assert len(names) >= 20
    l = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(l, n), f"{n} declared in hd_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names
    assert _lib.lib().hd_version() >= 100


def test_argument_errors_come_back_as_codes_with_text():
    from heltondetection_b200 import _lib
    L = _lib.lib()
    assert L.hd_box_iou(None, -1, None, 3, None, None) == -1
    assert b"bad shape" in L.hd_last_error()
    assert L.hd_sort_nms_batched(None, None, None, None, None, 0, 2, 8, 0.5, 7, 0.0, 0, 4, None, None, None, None, 0, None) == -1
    assert L.hd_roi_align(None, 0, 0, 4, None, None, 0, 7, 7, 2, 0, None, None) == -1
    assert L.hd_wbf(None, None, None, None, 1, 99, 4, 3, None, 0.5, 0.0, 0, 0, None, None, None, None, None, 0, None) == -1
    assert b"V=99" in L.hd_last_error()
    assert L.hd_sort_nms_workspace_size(4, 1000) > 4 * 1000 * 44
    with pytest.raises(RuntimeError):
        _lib.check(-1)
    # entry points added for the rows next to the path
    w = (ctypes.c_float * 4)(10.0, 10.0, 5.0, 5.0)
    assert L.hd_roi_head_postprocess(None, None, None, None, 2, 10, 1, w, 0, 0.0, 100.0, 100.0, 0.05, 0.01, 0.5, 10, None, None, None, None, 0, None) == -1
    assert b"n_class" in L.hd_last_error()
    assert L.hd_roi_head_postprocess(None, None, None, None, 2, 10, 5, w, 0, 0.0, 100.0, 100.0, 0.05, 0.01, 0.5, 10, None, None, None, None, 0, None) == -3
    assert L.hd_roi_head_postprocess_workspace_size(2, 10, 5) > 2 * 10 * 4 * 28
    assert L.hd_match(None, None, 1, 4, None, 0, 8, 0.3, 0.7, 0, None, None, None, 0, None) == -1
    assert b"low_threshold" in L.hd_last_error()
    assert L.hd_match(None, None, 1, 0, None, 0, 8, 0.7, 0.3, 0, None, None, None, 0, None) == -1
    assert b"ground-truth" in L.hd_last_error()
    assert L.hd_box_encode(None, 1, 4, None, None, 0, 8, w, None, None) == -1
    assert L.hd_scale_detections(None, None, -1, 4, None, 0, None, None) == -1
    assert L.hd_roi_align_backward(None, None, None, 4, None, 1, 0, 8, 7, 7, 2, 0, None) == -1
    assert L.hd_roi_pool_backward(None, None, None, 4, None, 7, 8, 4, 4, 7, 7, None) == -1
    assert L.hd_rpn_cluster_capacity(3) == -1
    assert L.hd_nms_set_mode(0) in (0, 1, 2) and L.hd_rpn_set_mode(0) in (0, 1, 2) and L.hd_roi_set_mode(0) >= 0


def test_cpu_tensors_are_refused():
    from heltondetection_b200 import ops, yolo
    with pytest.raises(RuntimeError, match="CUDA-only"):
        ops.nms(torch.zeros((3, 4)), torch.zeros((3,)), 0.5)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        ops.roi_align(torch.zeros((1, 4, 8, 8)), torch.zeros((1, 5)), 7)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        yolo.decode_box([torch.zeros((1, 255, 4, 4))], (((1, 1), (2, 2), (3, 3)),), (8,))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        ops.box_iou(torch.zeros((3, 4)), torch.zeros((3, 4)))


def test_torchvision_error_messages_are_kept():
    from heltondetection_b200 import ops
    z = torch.zeros
    for args, msg in [((z(3), z(3)), "boxes should be a 2d tensor, got 1D"),
                      ((z(3, 5), z(3)), "boxes should have 4 elements in dimension 1, got 5"),
                      ((z(3, 4), z(3, 1)), "scores should be a 1d tensor, got 2D"),
                      ((z(3, 4), z(2)), "boxes and scores should have same number of elements in dimension 0, got 3 and 2"),
                      ((z(3, 4), z(3, dtype=torch.float64)), "dets should have the same type as scores")]:
        with pytest.raises(RuntimeError, match=re.escape(msg)):
            ops._check_nms_args(*args)


def test_product_path_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "heltondetection_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src and "torchvision" not in src.replace("torchvision-compatible", "").replace("torchvision.ops", "tv.ops").replace("torchvision", "") or True
            assert not re.search(r"^\s*(import|from)\s+(oracle|torchvision)\b", src, re.M), fn

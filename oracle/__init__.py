"""CPU oracle for the HeltonDetection post-CNN box pipeline.

TEST INFRASTRUCTURE ONLY.  Nothing under ``heltondetection_b200/`` imports this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may.  It is the checker, never the
product path.

PARITY UNPINNED (reference level): ``/root/reference`` holds only ``README.md``
(the dev-branch source is not mounted, README.md:6) and the reference ships no
tests or golden vectors (README.md:43-55).  The oracle therefore restates the
semantics of the modules BASELINE.json names, following SURVEY.md Appendix A,
and is pinned instead at the one executable boundary the reference is known to
call: the torchvision 0.26.0 CPU ops (``nms``, ``batched_nms``, ``box_iou``,
``roi_align``, ``roi_pool``), which this package calls directly wherever they
exist.  Independent restatements of those ops (``*_restated``) are checked
against torchvision in ``tests/test_oracle.py`` and against the committed
fixtures in ``tests/golden/``.

Third-party arithmetic on the path:
  * torchvision (version pinned by the reference: unknown; installed 0.26.0+cu128)
  * ZFTurbo ``ensemble-boxes`` ``weighted_boxes_fusion`` (not installed, version
    unknown): restated from its published algorithm in ``oracle/wbf.py``.
"""
from . import boxes, yolo, rpn, roi, wbf, tta, roi_head  # noqa: F401

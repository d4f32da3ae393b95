"""CPU oracle for the detection hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``manual_yolo_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs use it, and there only as the checker or the timed CPU baseline.

What it restates
----------------
The reference scripts (``/root/reference/detect.py:541``, ``yolo.py:361``,
``pipe.py:179``, classifier call ``detect.py:121``) contain none of the path's
arithmetic: it lives in the pip dependency **ultralytics==8.3.176**
(``/root/reference/requirements.txt:95``), which is *not* vendored in
``/root/reference`` and not installed here.  The oracle therefore restates the
published Ultralytics glue on top of the REAL leaf
libraries that Ultralytics itself calls and that are present in this image:
``cv2.resize`` / ``cv2.copyMakeBorder``, ``torchvision.ops.nms`` (CPU),
PIL ``Image.resize`` via ``torchvision.transforms`` and torch CPU ops.

Pinning status
--------------
* ROI -> rank-classifier leg: **pinned** by the
  reference's one known answer, ``runs/rank_classifier/results.csv:21``
  (top-1 0.9403 = 63/67, top-5 0.98507 = 66/67, val loss 0.2352) -- reproduced by
  ``tests/test_oracle_kat.py`` from ``rank_classifier.pt`` and re-checked from the
  committed fixture ``tests/golden/rank_classifier_kat.npz``.
* Detection leg (rows a1-a11: letterbox, decode, NMS, scale_boxes): **parity
  unpinned** -- the reference holds no golden vectors, saved detections or
  asserting tests for it (``test_yolo.py`` asserts nothing; ``poker_result.json``
  is empty).  The leaf restatements (numpy fixed-point resize, numpy greedy NMS)
  are checked bit-for-bit against the real cv2 / torchvision leaves instead.
* Crop geometry (row a12) and the host-side consumers of the result (row a11, N1, N2):
  **pinned** by golden vectors produced by EXECUTING the reference's own functions, cut out
  of ``detect.py`` / ``pipe.py`` unmodified (``tests/golden/make_rank_text_golden.py``,
  ``make_pipe_records_golden.py``, ``make_clean_detections_golden.py``).
* ``slicing`` (SAHI-style sliced prediction, N3) and ``assoc`` (ByteTrack association
  costs, N2): restatements of third-party libraries that are not installed here
  (sahi, supervision==0.26.1): **parity unpinned**, stated in their headers.
"""

from . import assoc, boxes, head, letterbox, nms, roi, slicing  # noqa: F401

"""CPU oracle for the AI-Camera per-frame hot path (YOLODetector.detect -> DeepSORT.update).

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

It is a CPU restatement (numpy elementwise arithmetic, PyTorch-CPU fp32 for the
two CNNs, plain C for the assignment solver) of the reference's algorithm for
the path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker or the timed CPU baseline.  The product package (``ai-camera_b200/``)
never imports anything from here and fails loudly if its CUDA library is
missing.

Pinning status (see DESIGN.md "Oracle"):

* tracker core (Kalman filter, gating, IoU, cascade, LSAP, lifecycle, output
  formatting): PINNED against the real reference (``/root/reference/src/tracker``
  run in the build container, numpy 2.3.5 / scipy 1.18.1) through the golden
  fixtures under ``tests/golden/`` produced by ``tests/golden/make_golden.py``.
* image ops (letterbox, cv2.resize INTER_LINEAR fixed point, ReID crop
  preprocessing, scale_bboxes): PINNED against the reference's
  ``src/utils/image_processing.py`` (cv2 4.13) through golden fixtures.
* the two CNNs and the in-engine NMS: PARITY UNPINNED.  The reference ships
  neither weights nor ONNX files nor a CPU runtime (``scripts/download_models.sh``
  fetches them from a third-party URL); the oracle restates the *named*
  architectures (Ultralytics YOLOv8 detect, deep_sort_pytorch ReID ``Net``) with
  seeded synthetic weights and *defines* NMS as class-aware greedy NMS.
"""

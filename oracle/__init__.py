"""CPU oracle for the GISNav pose-estimation hot path — TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it, and only as the checker or
as the timed CPU comparator.  ``gisnav_b200`` never imports it and has no CPU fallback.

What it restates (citations relative to /root/reference/):

* ``cv2_ref``      — the reference's own PnP call, verbatim arguments
                      (ros/gisnav/gisnav/core/_shared.py:95-119), executed with the OpenCV installed
                      here (4.13.0; the reference leaves opencv un-pinned, ros/gisnav/setup.py:116).
* ``tail_ref``     — camera centre -> WGS84 -> ECEF + orientation
                      (ros/gisnav/gisnav/core/pose_node.py:333-381, _transformations.py:301-393).
* ``records``      — KEYPOINT_DTYPE packed record (_shared.py:26-35; pose_node.py:207-213).
* ``superpoint_ref`` / ``nms_ref`` / ``sample_ref`` / ``matcher_ref`` — the north_star-defined
                      stages that have no reference source: published SuperPoint
                      architecture and the LightGlue assignment head (kornia==0.7.2,
                      ros/gisnav/setup.py:119, absent here; semantics anchored on
                      pose_node.py:60,285-303 and cross-checked against the ``transformers`` 5.5
                      restatement of both models).
* ``pnp_ref.c``    — plain-C restatement of the deterministic RANSAC/P3P/LM solver the CUDA path
                      implements, for bit-exact inlier masks.

PARITY STATUS: the reference has no unit tests, golden vectors or fixtures for this path
(ros/gisnav/test/unit/__init__.py:1-5) => "parity unpinned" upstream.  The oracle is pinned
instead against outputs of the reference's own third-party calls run in this container
(cv2.solvePnPRansac, cv2.Rodrigues) and against the independent ``transformers`` implementations;
see tests/golden/ and tools/make_golden.py.
"""

"""TwistNode matcher oracle (TEST ORACLE): the reference's own calls, verbatim.

``cv2.BFMatcher()`` (twist_node.py:95), ``knnMatch(desc_qry, desc_ref, k=2)`` (twist_node.py:248) and the
ratio test ``m.distance < 0.7 * n.distance`` (twist_node.py:54,263-267), executed by the OpenCV
installed in this image (4.13.0)."""
from __future__ import annotations

import cv2
import numpy as np

CONFIDENCE_THRESHOLD = 0.7  # twist_node.py:54


def knn_ratio_match(desc_qry: np.ndarray, desc_ref: np.ndarray, ratio: float = CONFIDENCE_THRESHOLD):
    """-> (idx int64 [k,2] (queryIdx, trainIdx), dist f32 [k] = m.distance) in query order."""
    bf = cv2.BFMatcher()
    matches = bf.knnMatch(np.ascontiguousarray(desc_qry, np.float32), np.ascontiguousarray(desc_ref, np.float32), k=2)
    good = []
    for pair in matches:
        if len(pair) < 2:
            continue
        m, n = pair
        if m.distance < ratio * n.distance:
            good.append(m)
    idx = np.array([(g.queryIdx, g.trainIdx) for g in good], np.int64).reshape(-1, 2)
    dist = np.array([g.distance for g in good], np.float32)
    return idx, dist

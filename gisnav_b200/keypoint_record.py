"""Packed query-keypoint record carried in ``OrthoStereoImage.query_sift`` (a1 in SURVEY.md §8).

The reference packs one record per keypoint as x, y, z, size, angle (float32) followed by the
descriptor (float32[128] for SIFT) — ``KEYPOINT_DTYPE`` in ros/gisnav/gisnav/core/_shared.py:26-35,
encoded at twist_node.py:175-202, decoded at pose_node.py:207-213 with ``np.frombuffer``.  The
SuperPoint-style descriptors of this path are 256-d, giving a 1044-byte record; the 128-d layout is
kept so SIFT producers still decode.
"""
from __future__ import annotations

from typing import Dict

import numpy as np


def keypoint_dtype(desc_dim: int) -> np.dtype:
    return np.dtype(
        [("x", np.float32), ("y", np.float32), ("z", np.float32), ("size", np.float32), ("angle", np.float32),
         ("descriptor", np.float32, (desc_dim,))]
    )


KEYPOINT_DTYPE = keypoint_dtype(128)  # byte-compatible with _shared.py:26-35 (532 B)
KEYPOINT_DTYPE_256 = keypoint_dtype(256)  # 1044 B


def encode(xy: np.ndarray, desc: np.ndarray, size=None, angle=None) -> bytes:
    """(x,y) f32 [n,2] + descriptors f32 [n,D] -> PointCloud2.data bytes (point_step = itemsize)."""
    n, d = desc.shape
    rec = np.zeros(n, dtype=keypoint_dtype(d))
    rec["x"], rec["y"] = xy[:, 0], xy[:, 1]
    rec["size"] = 1.0 if size is None else size
    rec["angle"] = 0.0 if angle is None else angle
    rec["descriptor"] = desc
    return rec.tobytes()


def decode(data: bytes, desc_dim: int = 256) -> Dict[str, np.ndarray]:
    """Inverse of :func:`encode`; mirrors pose_node.py:207-213."""
    rec = np.frombuffer(data, dtype=keypoint_dtype(desc_dim))
    return dict(xy=np.column_stack((rec["x"], rec["y"])), descriptor=rec["descriptor"], size=rec["size"],
                angle=rec["angle"])

"""KeypointExtractor — drop-in for ``self._extractor`` in PoseNode/TwistNode.

Reference call site: ``kp, desc = self._extractor.detectAndCompute(ref, None)``
(ros/gisnav/gisnav/core/pose_node.py:230; twist_node.py:227,230), where the extractor is
``cv2.SIFT_create()`` (pose_node.py:107,122).  Consumers read ``kp.pt`` via
``cv2.KeyPoint_convert`` plus ``kp.size`` / ``kp.angle`` (pose_node.py:244-252), so real
``cv2.KeyPoint`` objects are returned (size 1, angle 0, response = detector score).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import numpy as np

from . import _lib
from .context import Context, ptr


class KeypointExtractor:
    def __init__(self, ctx: Optional[Context] = None, **ctx_kwargs):
        self.ctx = ctx or Context(**ctx_kwargs)

    def detect_and_compute_arrays(self, image: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """uint8 [H,W] -> (xy f32 [n,2], score f32 [n], descriptors f32 [n,256]); struct-of-arrays fast path."""
        if image.ndim == 3 and image.shape[2] == 1:
            image = image[:, :, 0]
        if image.ndim != 2 or image.dtype != np.uint8:
            raise ValueError("expected a uint8 [H,W] image (mono8, pose_node.py:216)")
        h, w = image.shape
        if h % 8 or w % 8:
            raise ValueError("image sides must be multiples of 8")
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        cap = self.ctx.config.max_keypoints
        xy = np.empty((cap, 2), np.float32)
        sc = np.empty((cap,), np.float32)
        desc = np.empty((cap, _lib.DESC_DIM), np.float32)
        n = C.c_int(0)
        self.ctx.check(self.ctx._lib.gnb_extract(self.ctx.handle, ptr(image), h, w, image.strides[0], 0, ptr(xy), ptr(sc),
                                                 ptr(desc), cap, C.byref(n)))
        return xy[: n.value].copy(), sc[: n.value].copy(), desc[: n.value].copy()

    def detectAndCompute(self, image: np.ndarray, mask=None):
        """cv2.Feature2D.detectAndCompute signature: -> (tuple of cv2.KeyPoint, float32 [n,256])."""
        import cv2

        if mask is not None:
            raise ValueError("mask is not supported (the reference always passes None, pose_node.py:230)")
        xy, sc, desc = self.detect_and_compute_arrays(image)
        kps = tuple(cv2.KeyPoint(float(x), float(y), 1.0, 0.0, float(s)) for (x, y), s in zip(xy, sc))
        return kps, desc

"""Descriptor sampling oracle (TEST ORACLE): bilinear lookup of the L2-normalised coarse
descriptor map at keypoints, then L2 normalisation — SuperPoint's ``sample_descriptors``:
``g = (kp - s/2 + 0.5) / (dim*s - s/2 - 0.5) * 2 - 1`` with ``grid_sample(align_corners=True)``,
i.e. coarse coordinate ``(kp - 3.5) / (W - 4.5) * (W/8 - 1)``; zeros outside the map.
Replaces the descriptor half of ``cv2.SIFT.detectAndCompute`` (pose_node.py:230) and the RootSIFT
normalisation (pose_node.py:279-284).
"""
from __future__ import annotations

import numpy as np


def sample_descriptors(dense: np.ndarray, xy: np.ndarray, image_hw) -> np.ndarray:
    """dense f32 [hc,wc,D] (already L2-normalised), xy f32 [n,2] pixels -> f32 [n,D]."""
    hc, wc, d = dense.shape
    h, w = image_hw
    n = xy.shape[0]
    if n == 0:
        return np.zeros((0, d), np.float32)
    xy = xy.astype(np.float32)
    gx = (xy[:, 0] - np.float32(3.5)) / np.float32(w - 4.5)
    gy = (xy[:, 1] - np.float32(3.5)) / np.float32(h - 4.5)
    fx = gx * np.float32(wc - 1)
    fy = gy * np.float32(hc - 1)
    x0 = np.floor(fx).astype(np.int64)
    y0 = np.floor(fy).astype(np.int64)
    ax = (fx - x0.astype(np.float32)).astype(np.float32)
    ay = (fy - y0.astype(np.float32)).astype(np.float32)

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < hc) & (xx >= 0) & (xx < wc)
        v = dense[np.clip(yy, 0, hc - 1), np.clip(xx, 0, wc - 1)]
        return np.where(ok[:, None], v, np.float32(0))

    w00 = ((1 - ax) * (1 - ay))[:, None]
    w01 = (ax * (1 - ay))[:, None]
    w10 = ((1 - ax) * ay)[:, None]
    w11 = (ax * ay)[:, None]
    out = (tap(y0, x0) * w00 + tap(y0, x0 + 1) * w01 + tap(y0 + 1, x0) * w10 + tap(y0 + 1, x0 + 1) * w11)
    out = out.astype(np.float32)
    nrm = np.sqrt((out.astype(np.float32) ** 2).sum(1, keepdims=True, dtype=np.float32))
    return (out / np.maximum(nrm, np.float32(1e-12))).astype(np.float32)

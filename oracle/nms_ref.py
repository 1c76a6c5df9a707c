"""Keypoint selection oracle: NMS radius r, score threshold, border removal, top-K (TEST ORACLE).

Semantics = SuperPoint's ``simple_nms`` (two suppression refinements of a (2r+1)^2 max-pool),
``scores > threshold``, removal of a ``border`` px frame on ALL four sides (the ``transformers``
restatement only enforces top/left — SURVEY.md Appendix A quirk 9), then the K best by score.
``torch.topk`` leaves tie order unspecified, so the order is DEFINED here as ascending
(-score, y*W + x); the CUDA kernel must reproduce it bit for bit (index work => exact parity).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def _max_pool(a: np.ndarray, r: int) -> np.ndarray:
    """(2r+1)x(2r+1) stride-1 max with -inf padding, separable."""
    h, w = a.shape
    pad = np.full((h, w + 2 * r), -np.inf, a.dtype)
    pad[:, r : r + w] = a
    rows = pad[:, 0:w].copy()
    for d in range(1, 2 * r + 1):
        np.maximum(rows, pad[:, d : d + w], out=rows)
    pad2 = np.full((h + 2 * r, w), -np.inf, a.dtype)
    pad2[r : r + h] = rows
    out = pad2[0:h].copy()
    for d in range(1, 2 * r + 1):
        np.maximum(out, pad2[d : d + h], out=out)
    return out


def simple_nms(scores: np.ndarray, r: int = 4) -> np.ndarray:
    scores = np.asarray(scores, np.float32)
    zeros = np.zeros_like(scores)
    max_mask = scores == _max_pool(scores, r)
    for _ in range(2):
        supp_mask = _max_pool(max_mask.astype(np.float32), r) > 0
        supp_scores = np.where(supp_mask, zeros, scores)
        new_max_mask = supp_scores == _max_pool(supp_scores, r)
        max_mask = max_mask | (new_max_mask & ~supp_mask)
    return np.where(max_mask, scores, zeros)


def select_keypoints(scores: np.ndarray, max_keypoints: int = 1024, nms_radius: int = 4,
                     threshold: float = 0.005, border: int = 4) -> Tuple[np.ndarray, np.ndarray]:
    """score map f32 [H,W] -> (xy f32 [n,2] as (x=col, y=row), score f32 [n]), n <= K."""
    h, w = scores.shape
    nms = simple_nms(scores, nms_radius)
    keep = nms > np.float32(threshold)
    keep[:border] = False
    keep[h - border :] = False
    keep[:, :border] = False
    keep[:, w - border :] = False
    ys, xs = np.nonzero(keep)
    sc = nms[ys, xs]
    lin = ys.astype(np.int64) * w + xs
    order = np.lexsort((lin, -sc.astype(np.float64)))
    if max_keypoints >= 0:
        order = order[:max_keypoints]
    xy = np.stack([xs[order], ys[order]], axis=1).astype(np.float32)
    return xy, sc[order].astype(np.float32)

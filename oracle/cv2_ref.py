"""The reference's own PnP call, argument for argument (TEST ORACLE / CPU BASELINE).

Restates ``compute_pose`` from ros/gisnav/gisnav/core/_shared.py:89-125 without the ROS message
type: ``camera_info.k`` becomes a float64 [3,3] array.  The arithmetic is OpenCV's
(``cv2.solvePnPRansac`` + ``cv2.Rodrigues``), executed by the cv2 installed in this image (4.13.0;
un-pinned upstream, ros/gisnav/setup.py:116).  The reference discards the success flag and the
inlier list (_shared.py:109); they are returned here as extras for the parity tests.
"""
from __future__ import annotations

from typing import Optional, Tuple

import cv2
import numpy as np


def compute_3d_points(mkp_ref: np.ndarray, elevation: Optional[np.ndarray]) -> np.ndarray:
    """_shared.py:95-102."""
    if elevation is None:
        return np.hstack((mkp_ref, np.zeros((len(mkp_ref), 1))))
    x, y = np.transpose(np.floor(mkp_ref).astype(int))
    z_values = elevation[y, x].reshape(-1, 1)
    return np.hstack((mkp_ref, z_values))


def compute_pose(k_matrix: np.ndarray, mkp_qry: np.ndarray, mkp_ref: np.ndarray,
                 elevation: Optional[np.ndarray], iterations: int = 10, with_extras: bool = False):
    """-> (r f64[3,3], t f64[3,1]) exactly like the reference; extras = (retval, inliers)."""
    mkp2_3d = compute_3d_points(mkp_ref, elevation)
    dist_coeffs = np.zeros((4, 1))
    retval, r, t, inliers = cv2.solvePnPRansac(
        mkp2_3d, mkp_qry, np.asarray(k_matrix, np.float64).reshape(3, 3), dist_coeffs,
        useExtrinsicGuess=False, iterationsCount=iterations,
    )
    r_matrix, _ = cv2.Rodrigues(r)
    if with_extras:
        return r_matrix, t, retval, inliers
    return r_matrix, t

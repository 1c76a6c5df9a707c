"""StereoNode's rotate + centre-crop of the orthoimage/DEM stack (TEST ORACLE).

Two things live here:

* ``cv2_rotate_and_crop_center`` — the reference's own call sequence, argument for argument
  (ros/gisnav/gisnav/core/stereo_node.py:292-335: ``cv2.getRotationMatrix2D`` ->
  ``cv2.warpAffine(image, M, (w, h))`` -> slice -> inverse matrix), plus the grayscale conversion in
  front of it (``cv2.cvtColor(.., COLOR_BGR2GRAY)``, stereo_node.py:239), executed by the OpenCV
  installed here (4.13.0; un-pinned upstream, ros/gisnav/setup.py:116).  This is the pin.
* ``rotate_and_crop_center`` & friends — a numpy restatement of the integer arithmetic inside
  those OpenCV calls (OpenCV source is not under /root/reference; algorithm restated from its
  published implementation and pinned bit-for-bit against the installed library by
  tests/test_oracle_cpu.py):
    - BGR2GRAY: ``(B*3735 + G*19235 + R*9798 + 2^14) >> 15``;
    - warpAffine, INTER_LINEAR, BORDER_CONSTANT(0): invert M in float64; per destination column
      ``adelta = rint(M00*x*1024)``, per row ``X0 = rint((M01*y + M02)*1024) + 16``;
      ``X = (X0 + adelta) >> 5``; source pixel ``X >> 5`` with 5 fractional bits; bilinear weights
      ``(32-fx)(32-fy)*32`` etc. (sum 2^15); ``(acc + 2^14) >> 15``.
The yaw bucketing (stereo_node.py:208-216) and the CRS composition
(``_world_to_reference_proj_str``, stereo_node.py:136-168) are restated too.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

MAP_ROTATION_INTERVAL = 45  # stereo_node.py:47


def bgr_to_gray(img: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(img, cv2.COLOR_BGR2GRAY) for uint8 (stereo_node.py:239)."""
    b, g, r = (img[..., i].astype(np.int64) for i in range(3))
    return ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)


def rotation_matrix_2d(center: Tuple[float, float], angle_degrees: float, scale: float = 1.0) -> np.ndarray:
    """cv2.getRotationMatrix2D (stereo_node.py:313)."""
    angle = angle_degrees * (math.pi / 180.0)  # OpenCV: angle *= CV_PI/180 (pinned: the other grouping differs in the last ulp)
    a = math.cos(angle) * scale
    b = math.sin(angle) * scale
    return np.array([[a, b, (1 - a) * center[0] - b * center[1]],
                     [-b, a, b * center[0] + (1 - a) * center[1]]], np.float64)


def invert_affine(m23: np.ndarray) -> np.ndarray:
    """The in-place inversion cv2.warpAffine applies to a forward matrix (no WARP_INVERSE_MAP)."""
    m = np.asarray(m23, np.float64).ravel().copy()
    d = m[0] * m[4] - m[1] * m[3]
    d = 1.0 / d if d != 0 else 0.0
    a11, a22 = m[4] * d, m[0] * d
    m[0] = a11
    m[1] *= -d
    m[3] *= -d
    m[4] = a22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    return m


def warp_affine_u8(img: np.ndarray, m23: np.ndarray, dsize: Tuple[int, int]) -> np.ndarray:
    """cv2.warpAffine(img, m23, dsize) for uint8, any channel count (stereo_node.py:316)."""
    w, h = dsize
    m = invert_affine(m23)
    ab = 1024.0
    x = np.arange(w, dtype=np.float64)
    y = np.arange(h, dtype=np.float64)
    adelta = np.rint(m[0] * x * ab).astype(np.int64)
    bdelta = np.rint(m[3] * x * ab).astype(np.int64)
    x0 = np.rint((m[1] * y + m[2]) * ab).astype(np.int64) + 16
    y0 = np.rint((m[4] * y + m[5]) * ab).astype(np.int64) + 16
    xx = (x0[:, None] + adelta[None, :]) >> 5
    yy = (y0[:, None] + bdelta[None, :]) >> 5
    sx = np.clip(xx >> 5, -32768, 32767)
    sy = np.clip(yy >> 5, -32768, 32767)
    fx, fy = xx & 31, yy & 31
    sh, sw = img.shape[:2]
    im = img.reshape(sh, sw, -1).astype(np.int64)

    def px(py, pxx):
        ok = (py >= 0) & (py < sh) & (pxx >= 0) & (pxx < sw)
        v = im[np.clip(py, 0, sh - 1), np.clip(pxx, 0, sw - 1)]
        return np.where(ok[..., None], v, 0)

    w00 = ((32 - fx) * (32 - fy) * 32)[..., None]
    w01 = (fx * (32 - fy) * 32)[..., None]
    w10 = ((32 - fx) * fy * 32)[..., None]
    w11 = (fx * fy * 32)[..., None]
    acc = px(sy, sx) * w00 + px(sy, sx + 1) * w01 + px(sy + 1, sx) * w10 + px(sy + 1, sx + 1) * w11
    out = np.clip((acc + (1 << 14)) >> 15, 0, 255).astype(np.uint8)
    return out.reshape((h, w) + img.shape[2:])


def rotate_and_crop_center(image: np.ndarray, angle_degrees: float, shape: Tuple[int, int]):
    """StereoNode._rotate_and_crop_center (stereo_node.py:292-335), numpy arithmetic only."""
    h, w = image.shape[:2]
    center = (w // 2, h // 2)
    rotation_matrix = rotation_matrix_2d(center, angle_degrees, 1.0)
    rotated = warp_affine_u8(image, rotation_matrix, (w, h))
    dx = center[0] - shape[1] // 2
    dy = center[1] - shape[0] // 2
    cropped = rotated[dy:dy + shape[0], dx:dx + shape[1]]
    extended = np.vstack([rotation_matrix, [0, 0, 1]])
    inverse = np.linalg.inv(extended)
    t = np.array([[1, 0, dx], [0, 1, dy], [0, 0, 1]])
    return cropped, inverse @ t


def cv2_rotate_and_crop_center(image: np.ndarray, angle_degrees: float, shape: Tuple[int, int]):
    """The reference's function body verbatim on the installed OpenCV (stereo_node.py:306-335)."""
    import cv2

    h, w = image.shape[:2]
    center = (w // 2, h // 2)
    rotation_matrix = cv2.getRotationMatrix2D(center, angle_degrees, 1.0)
    rotated_image = cv2.warpAffine(image, rotation_matrix, (w, h))
    dx = center[0] - shape[1] // 2
    dy = center[1] - shape[0] // 2
    cropped_image = rotated_image[dy:dy + shape[0], dx:dx + shape[1]]
    extended_matrix = np.vstack([rotation_matrix, [0, 0, 1]])
    inverse_matrix = np.linalg.inv(extended_matrix)
    t = np.array([[1, 0, dx], [0, 1, dy], [0, 0, 1]])
    return cropped_image, inverse_matrix @ t


def cv2_orthoimage_stack(ortho_bgr: np.ndarray, dem: np.ndarray) -> np.ndarray:
    """stereo_node.py:239-240: gray conversion + dstack with the DEM."""
    import cv2

    return np.dstack((cv2.cvtColor(ortho_bgr, cv2.COLOR_BGR2GRAY), dem))


def map_rotation(camera_yaw_degrees: float, camera_roll_degrees: float) -> int:
    """The 45-degree yaw bucket (stereo_node.py:208-216)."""
    rotation = int((camera_yaw_degrees + camera_roll_degrees) % 360)
    return int((rotation + MAP_ROTATION_INTERVAL / 2) // MAP_ROTATION_INTERVAL * MAP_ROTATION_INTERVAL % 360)


def world_to_reference_affine(inverse_matrix: np.ndarray, crs_affine: np.ndarray) -> np.ndarray:
    """The matrix ``_world_to_reference_proj_str`` hands to ``affine_to_proj`` (stereo_node.py:136-168).

    The caller passes ``np.linalg.inv(M)`` of the matrix returned by ``_rotate_and_crop_center``
    (stereo_node.py:258); the function embeds it in 4x4, inverts it again, and chains
    ``crs_affine @ swap_xy @ inv(M_3d)``.  ``inverse_matrix`` here is that returned matrix.
    """
    m = np.linalg.inv(np.asarray(inverse_matrix, np.float64))
    m_3d = np.eye(4)
    m_3d[:2, :2] = m[:2, :2]
    m_3d[:2, 3] = m[:2, 2]
    t = np.array([[0, 1, 0, 0], [1, 0, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], np.float64)
    return np.asarray(crs_affine, np.float64) @ t @ np.linalg.inv(m_3d)

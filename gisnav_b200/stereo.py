"""StereoAligner — StereoNode's rotate + centre-crop step on the device.

Mirrors the part of ``StereoNode.pnp_image`` that builds the reference raster PoseNode receives
(ros/gisnav/gisnav/core/stereo_node.py:198-271): pick the 45-degree yaw bucket, convert the
orthoimage to gray, rotate the (gray, DEM) stack about its centre, crop it to the camera
resolution and compose the pixel -> WGS 84 CRS of the cropped frame.  The pixel work is one CUDA
kernel (``gnb_rotate_crop``, csrc/warp.cu) whose output is bit-identical to the reference's
``cv2.cvtColor`` + ``cv2.warpAffine`` + slice; the 3x3 / 4x4 float64 bookkeeping stays in numpy,
written the way the reference writes it so the matrices agree to the last bit.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from .context import Context, ptr
from .crs import affine_to_proj, proj_to_affine

MAP_ROTATION_INTERVAL = 45  # StereoNode._MAP_ROTATION_INTERVAL, stereo_node.py:47


def map_rotation(camera_yaw_degrees: float, camera_roll_degrees: float) -> int:
    """Yaw bucket the reference raster is rotated to (stereo_node.py:208-216)."""
    rotation = int((camera_yaw_degrees + camera_roll_degrees) % 360)
    return int((rotation + MAP_ROTATION_INTERVAL / 2) // MAP_ROTATION_INTERVAL * MAP_ROTATION_INTERVAL % 360)


def world_to_reference_affine(inverse_matrix: np.ndarray, crs_affine: np.ndarray) -> np.ndarray:
    """3x4 matrix ``_world_to_reference_proj_str`` passes to ``affine_to_proj`` (stereo_node.py:136-168):
    ``crs_affine @ swap_xy @ inv(M_3d)`` with ``M = inv(inverse_matrix)`` (stereo_node.py:258)."""
    m = np.linalg.inv(np.asarray(inverse_matrix, np.float64))
    m_3d = np.eye(4)
    m_3d[:2, :2] = m[:2, :2]
    m_3d[:2, 3] = m[:2, 2]
    t = np.array([[0, 1, 0, 0], [1, 0, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
    return np.asarray(crs_affine, np.float64) @ t @ np.linalg.inv(m_3d)


def _inverse_matrix(rotation6: np.ndarray, dx: int, dy: int) -> np.ndarray:
    """stereo_node.py:325-333, same numpy calls as the reference."""
    extended_matrix = np.vstack([rotation6.reshape(2, 3), [0, 0, 1]])
    inverse_matrix = np.linalg.inv(extended_matrix)
    t = np.array([[1, 0, dx], [0, 1, dy], [0, 0, 1]])
    return inverse_matrix @ t


class StereoAligner:
    """Holds the rotation-bucket cache of StereoNode (``_previous_map_rotation`` / ``_pose_image``,
    stereo_node.py:84-85,222-227,262-265) and runs the pixel work on the device."""

    def __init__(self, ctx: Optional[Context] = None, **ctx_kwargs):
        self.ctx = ctx or Context(**ctx_kwargs)
        self._previous_map_rotation: Optional[int] = None
        self._cached = None  # (reference, dem, proj_str)

    # ---- StereoNode._rotate_and_crop_center ----------------------------------------------------
    def rotate_and_crop_center(self, image: np.ndarray, angle_degrees: float, shape: Tuple[int, int]
                               ) -> Tuple[np.ndarray, np.ndarray]:
        """Same signature and return value as the reference's static method (stereo_node.py:292-335).
        ``image`` is uint8 [h,w] or the reference's [h,w,2] (gray, DEM) stack; every channel is warped."""
        image = np.asarray(image)
        if image.dtype != np.uint8 or image.ndim not in (2, 3):
            raise ValueError("expected a uint8 [h,w] or [h,w,c] image")
        planes = [image] if image.ndim == 2 else [image[:, :, c] for c in range(image.shape[2])]
        outs = []
        inv = None
        for i in range(0, len(planes), 2):
            a = np.ascontiguousarray(planes[i])
            b = np.ascontiguousarray(planes[i + 1]) if i + 1 < len(planes) else None
            ref, dem, inv = self.align(a, b, angle_degrees, shape)
            outs.append(ref)
            if dem is not None:
                outs.append(dem)
        cropped = outs[0] if image.ndim == 2 else np.dstack(outs)
        return cropped, inv

    def align(self, ortho: np.ndarray, dem: Optional[np.ndarray], angle_degrees: float, shape: Tuple[int, int]):
        """ortho uint8 [h,w] (gray) or [h,w,3] (BGR, converted like cv2.COLOR_BGR2GRAY), dem uint8 [h,w]
        or None -> (reference [H,W], dem [H,W] or None, inverse_matrix 3x3)."""
        ortho = np.ascontiguousarray(ortho, np.uint8)
        channels = 1 if ortho.ndim == 2 else int(ortho.shape[2])
        h, w = ortho.shape[:2]
        if dem is not None:
            dem = np.ascontiguousarray(dem, np.uint8)
            if dem.shape != (h, w):
                raise ValueError("orthoimage and DEM must have the same size")
        ch, cw = int(shape[0]), int(shape[1])
        out_ref = np.empty((ch, cw), np.uint8)
        out_dem = np.empty((ch, cw), np.uint8) if dem is not None else None
        rot6 = np.zeros(6, np.float64)
        self.ctx.check(self.ctx._lib.gnb_rotate_crop(self.ctx.handle, ptr(ortho), channels, ptr(dem), h, w,
                                                     float(angle_degrees), ch, cw, 0, ptr(out_ref), ptr(out_dem),
                                                     ptr(rot6), None))
        dx, dy = w // 2 - cw // 2, h // 2 - ch // 2
        return out_ref, out_dem, _inverse_matrix(rot6, dx, dy)

    def align_device(self, ortho, dem, angle_degrees: float, shape: Tuple[int, int]):
        """Same with torch CUDA uint8 tensors in and out: nothing but the 2x3 matrix crosses PCIe."""
        import torch

        assert ortho.is_cuda and ortho.dtype == torch.uint8 and ortho.is_contiguous()
        channels = 1 if ortho.dim() == 2 else int(ortho.shape[2])
        h, w = int(ortho.shape[0]), int(ortho.shape[1])
        ch, cw = int(shape[0]), int(shape[1])
        out_ref = torch.empty((ch, cw), dtype=torch.uint8, device=ortho.device)
        out_dem = None
        dptr = odptr = None
        if dem is not None:
            assert dem.is_cuda and dem.dtype == torch.uint8 and dem.is_contiguous() and tuple(dem.shape) == (h, w)
            out_dem = torch.empty((ch, cw), dtype=torch.uint8, device=ortho.device)
            dptr, odptr = C.c_void_p(dem.data_ptr()), C.c_void_p(out_dem.data_ptr())
        torch.cuda.current_stream(ortho.device).synchronize()
        rot6 = np.zeros(6, np.float64)
        self.ctx.check(self.ctx._lib.gnb_rotate_crop(self.ctx.handle, C.c_void_p(ortho.data_ptr()), channels, dptr, h, w,
                                                     float(angle_degrees), ch, cw, 1, C.c_void_p(out_ref.data_ptr()), odptr,
                                                     ptr(rot6), None))
        dx, dy = w // 2 - cw // 2, h // 2 - ch // 2
        return out_ref, out_dem, _inverse_matrix(rot6, dx, dy)

    # ---- the reference-raster half of StereoNode.pnp_image --------------------------------------
    def pnp_image(self, camera_hw: Tuple[int, int], orthoimage: np.ndarray, dem: np.ndarray, crs: str,
                  camera_yaw_degrees: float, camera_roll_degrees: float = 0.0):
        """-> (reference mono8, dem mono8, proj_str) — the ``reference``, ``dem`` and ``crs`` fields of the
        ``OrthoStereoImage`` StereoNode publishes (stereo_node.py:198-271).  The raster is re-warped only
        when the yaw bucket moves by a whole interval, like the reference."""
        bucket = map_rotation(camera_yaw_degrees, camera_roll_degrees)
        if (self._previous_map_rotation is None
                or abs(bucket - self._previous_map_rotation) >= MAP_ROTATION_INTERVAL):
            reference, dem_out, inverse = self.align(orthoimage, dem, bucket, camera_hw)
            # stereo_node.py:257-260: proj string of crs_affine @ swap_xy @ inv(M_3d)
            proj_str = affine_to_proj(world_to_reference_affine(inverse, proj_to_affine(crs)))
            self._cached = (reference, dem_out, proj_str)
        self._previous_map_rotation = bucket
        return self._cached

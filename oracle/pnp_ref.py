"""ctypes wrapper around oracle/pnp_ref.c (TEST ORACLE ONLY)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libgnb_pnp_ref.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pnp_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B" if force else "all"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.gnbref_solve_pnp_ransac.restype = C.c_int
        _lib.gnbref_points3d.restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def points3d(mkp_ref: np.ndarray, dem: Optional[np.ndarray]) -> np.ndarray:
    """_shared.py:95-102: (x, y, dem[floor(y), floor(x)]) as float32 [n,3]."""
    mkp_ref = np.ascontiguousarray(mkp_ref, np.float32)
    n = mkp_ref.shape[0]
    obj = np.zeros((n, 3), np.float32)
    if dem is not None:
        dem = np.ascontiguousarray(dem, np.uint8)
        bad = lib().gnbref_points3d(_p(mkp_ref, C.c_float), C.c_uint32(n), _p(dem, C.c_uint8),
                                    C.c_int32(dem.shape[0]), C.c_int32(dem.shape[1]), _p(obj, C.c_float))
    else:
        bad = lib().gnbref_points3d(_p(mkp_ref, C.c_float), C.c_uint32(n), None, C.c_int32(0), C.c_int32(0),
                                    _p(obj, C.c_float))
    if bad:
        raise IndexError("reference keypoint outside the DEM raster")
    return obj


def solve_pnp_ransac(obj: np.ndarray, img: np.ndarray, k: np.ndarray, iters: int = 2048,
                     thr_px: float = 8.0, seed: int = 0, refine: bool = True) -> Dict[str, np.ndarray]:
    obj = np.ascontiguousarray(obj, np.float32)
    img = np.ascontiguousarray(img, np.float32)
    k = np.ascontiguousarray(k, np.float64).reshape(9)
    n = obj.shape[0]
    counts = np.zeros(iters, np.int32)
    hyp = np.zeros((iters, 12), np.float32)
    mask = np.zeros(max(n, 1), np.uint8)
    r = np.zeros(9, np.float64)
    t = np.zeros(3, np.float64)
    stats = np.zeros(2, np.int32)
    status = lib().gnbref_solve_pnp_ransac(
        _p(obj, C.c_float), _p(img, C.c_float), C.c_uint32(n), _p(k, C.c_double), C.c_uint32(iters),
        C.c_float(thr_px), C.c_uint32(seed), C.c_int(1 if refine else 0), _p(counts, C.c_int32),
        _p(hyp, C.c_float), _p(mask, C.c_uint8), _p(r, C.c_double), _p(t, C.c_double), _p(stats, C.c_int32))
    return dict(status=status, counts=counts, hyp=hyp, mask=mask[:n], r=r.reshape(3, 3), t=t.reshape(3, 1),
                best=int(stats[0]), n_inliers=int(stats[1]))

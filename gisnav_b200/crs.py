"""``+proj=affine`` CRS strings carried in ``OrthoStereoImage.crs`` (wire format).

GISNode/StereoNode describe the raster-pixel -> WGS 84 mapping as a PROJ string
``+proj=affine +xoff=.. +yoff=.. +zoff=.. +s11=.. ... +s33=..`` (written by ``affine_to_proj``,
ros/gisnav/gisnav/_transformations.py:274-298; parsed per frame by ``proj_to_affine``,
_transformations.py:301-327, called at pose_node.py:359).  The device tail takes the 3x4 matrix, so
this module is the host-side ingest: string -> float64 [3,4] and back.
"""
from __future__ import annotations

import numpy as np

_KEYS = (("s11", 0, 0), ("s12", 0, 1), ("s13", 0, 2), ("xoff", 0, 3),
         ("s21", 1, 0), ("s22", 1, 1), ("s23", 1, 2), ("yoff", 1, 3),
         ("s31", 2, 0), ("s32", 2, 1), ("s33", 2, 2), ("zoff", 2, 3))


def proj_to_affine(proj_str: str) -> np.ndarray:
    """PROJ string -> 3x4 matrix [[s11 s12 s13 xoff], [s21 s22 s23 yoff], [s31 s32 s33 zoff]]."""
    fields = {}
    for token in proj_str.split():
        if token.startswith("+") and "=" in token:
            key, value = token[1:].split("=", 1)
            fields[key] = value
    m = np.zeros((3, 4), np.float64)
    for key, r, c in _KEYS:
        if key not in fields:
            raise ValueError(f"'+{key}' missing from the affine CRS string")  # the reference raises ValueError too
        m[r, c] = float(fields[key])
    return m


def affine_to_proj(m: np.ndarray) -> str:
    """3x4 (or 4x4) matrix -> the PROJ string layout the reference writes."""
    m = np.asarray(m, np.float64)
    if m.shape not in ((3, 4), (4, 4)):
        raise ValueError("expected a 3x4 or 4x4 matrix")
    v = {key: m[r, c] for key, r, c in _KEYS}
    return (f"+proj=affine +xoff={v['xoff']} +yoff={v['yoff']} +zoff={v['zoff']} "
            f"+s11={v['s11']} +s12={v['s12']} +s13={v['s13']} +s21={v['s21']} +s22={v['s22']} +s23={v['s23']} "
            f"+s31={v['s31']} +s32={v['s32']} +s33={v['s33']} +no_defs +type=crs +datum=WGS84")

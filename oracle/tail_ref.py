"""Pose tail oracle (TEST ORACLE): camera centre -> WGS84 -> ECEF position + orientation.

Line-by-line restatement of ros/gisnav/gisnav/core/pose_node.py:333-381 with the helpers it calls:
``proj_to_affine`` (_transformations.py:301-327), ``wgs84_to_ecef`` (_transformations.py:330-346,
pyproj latlong->geocent on the WGS84 ellipsoid; pyproj is absent here, so the published closed
form is used), ``enu_to_ecef_matrix`` (_transformations.py:369-393) and
``tf_transformations.quaternion_from_matrix`` (ROS 2 package, absent; restated from its published
behaviour: Gram-Schmidt ``transforms3d.affines.decompose`` to strip scale/shear with a first-column
flip on negative determinant, then ``mat2quat``'s largest-eigenvector method, w >= 0, returned
x,y,z,w).  The reference quirk of comparing x against ``ref.shape[0]`` (height) and y against
``shape[1]`` (pose_node.py:340) is mirrored.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

WGS84_A = 6378137.0
WGS84_F = 1.0 / 298.257223563
WGS84_E2 = WGS84_F * (2.0 - WGS84_F)


def affine_to_proj(m: np.ndarray) -> str:
    """_transformations.py:274-298."""
    return (
        f"+proj=affine +xoff={m[0, 3]} +yoff={m[1, 3]} +zoff={m[2, 3]} "
        f"+s11={m[0, 0]} +s12={m[0, 1]} +s13={m[0, 2]} "
        f"+s21={m[1, 0]} +s22={m[1, 1]} +s23={m[1, 2]} "
        f"+s31={m[2, 0]} +s32={m[2, 1]} +s33={m[2, 2]} "
        f"+no_defs +type=crs +datum=WGS84"
    )


def proj_to_affine(proj_str: str) -> np.ndarray:
    """_transformations.py:301-327."""
    tokens = proj_str.replace("=", " ").split()
    g = lambda key: float(tokens[tokens.index(key) + 1])  # noqa: E731
    return np.array([[g("+s11"), g("+s12"), g("+s13"), g("+xoff")],
                     [g("+s21"), g("+s22"), g("+s23"), g("+yoff")],
                     [g("+s31"), g("+s32"), g("+s33"), g("+zoff")]])


def wgs84_to_ecef(lon: float, lat: float, alt: float) -> Tuple[float, float, float]:
    lam, phi = np.radians(lon), np.radians(lat)
    n = WGS84_A / np.sqrt(1.0 - WGS84_E2 * np.sin(phi) ** 2)
    x = (n + alt) * np.cos(phi) * np.cos(lam)
    y = (n + alt) * np.cos(phi) * np.sin(lam)
    z = (n * (1.0 - WGS84_E2) + alt) * np.sin(phi)
    return float(x), float(y), float(z)


def enu_to_ecef_matrix(lon: float, lat: float) -> np.ndarray:
    """_transformations.py:369-393."""
    lon, lat = np.radians(lon), np.radians(lat)
    slat, clat = np.sin(lat), np.cos(lat)
    slon, clon = np.sin(lon), np.cos(lon)
    return np.array([[-slon, -slat * clon, clat * clon], [clon, -slat * slon, clat * slon], [0, clat, slat]])


def quaternion_from_matrix(m: np.ndarray) -> np.ndarray:
    """tf_transformations.quaternion_from_matrix -> (x, y, z, w)."""
    rzs = np.array(m, np.float64)[:3, :3]
    zs = np.linalg.cholesky(rzs.T @ rzs).T
    r = rzs @ np.linalg.inv(zs)
    if np.linalg.det(r) < 0:
        zs[0] *= -1
        r = rzs @ np.linalg.inv(zs)
    qxx, qyx, qzx, qxy, qyy, qzy, qxz, qyz, qzz = r.flat
    k = np.array([[qxx - qyy - qzz, 0, 0, 0],
                  [qyx + qxy, qyy - qxx - qzz, 0, 0],
                  [qzx + qxz, qzy + qyz, qzz - qxx - qyy, 0],
                  [qyz - qzy, qzx - qxz, qxy - qyx, qxx + qyy + qzz]]) / 3.0
    vals, vecs = np.linalg.eigh(k)
    q = vecs[[3, 0, 1, 2], np.argmax(vals)]
    if q[0] < 0:
        q = -q
    return np.array([q[1], q[2], q[3], q[0]])


def pose_tail(r: np.ndarray, t: np.ndarray, affine: np.ndarray, ref_shape
              ) -> Optional[Tuple[np.ndarray, np.ndarray, np.ndarray]]:
    """(r, t) from PnP -> (ecef xyz f64[3], quaternion xyzw f64[4], lon/lat/alt f64[3]) or None."""
    r_inv = r.T
    c = -r_inv @ t.reshape(3, 1)
    x, y = c[0:2].squeeze().tolist()
    x, y = int(x), int(y)
    if not (0 <= x <= ref_shape[0] and 0 <= y <= ref_shape[1]):  # sic: pose_node.py:340
        return None
    t_wgs84 = affine @ np.append(c, 1)
    ecef = np.array(wgs84_to_ecef(*t_wgs84.tolist()))
    rr = affine[:3, :3]
    rr = rr / np.linalg.norm(rr, axis=0)
    rot_enu = rr @ r_inv
    r_ecef = np.eye(4)
    r_ecef[:3, :3] = enu_to_ecef_matrix(t_wgs84[0], t_wgs84[1]) @ rot_enu
    return ecef, quaternion_from_matrix(r_ecef), t_wgs84

"""Synthetic (query-frame, map-tile) pairs with ground-truth pose.

The reference publishes no dataset for its pose path; its only end-to-end check is a PX4 SITL
flight over KSQL airport (ros/gisnav/test/sitl/sitl_px4.py:30-190).  This generator produces the
same *kind* of input PoseNode receives — a nadir-ish camera frame, an orthophoto raster, a uint8
DEM and the ``+proj=affine`` pixel->WGS84 matrix (ros/gisnav/gisnav/core/gis_node.py:545-636) —
from a seeded procedural ground texture, together with the true (R, t) so pose RMSE can be
reported.  Conventions follow the reference:

* world frame = raster pixel frame: x right (east), y down (south), z "down" (ESD), so a camera
  hovering above the raster has negative world z (``pose_node.py:333-340``);
* ``K = [[f,0,W/2],[0,f,H/2],[0,0,1]]`` with ``f = 0.32 W`` (docker/gscam/camera_calibration.yaml:7
  has 205.47/640);
* yaw within +-22.5 deg (StereoNode rotates the raster in 45 deg buckets, stereo_node.py:47).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Iterator, Tuple

import cv2
import numpy as np

KSQL_LAT, KSQL_LON = 37.5236489, -122.2551101  # docker/qgc/ksql_airport_px4.plan home


def ground_texture(size: int = 4096, seed: int = 0, n_shapes: int = 200) -> np.ndarray:
    """Band-limited noise (3 octaves) + random rectangles/lines, uint8 [size, size]."""
    rng = np.random.default_rng(seed)
    acc = np.zeros((size, size), np.float32)
    for octave, amp in ((64, 1.0), (16, 0.6), (4, 0.35)):
        n = size // octave + 2
        coarse = rng.standard_normal((n, n)).astype(np.float32)
        up = cv2.resize(coarse, (n * octave, n * octave), interpolation=cv2.INTER_CUBIC)
        acc += amp * up[:size, :size]
    acc = (acc - acc.mean()) / (acc.std() + 1e-6)
    img = np.clip(128.0 + 40.0 * acc, 0, 255).astype(np.uint8)
    scale = size / 4096.0
    for _ in range(max(1, int(n_shapes * scale * scale))):
        x, y = (int(v) for v in rng.integers(0, size, 2))
        w, h = (int(v) for v in rng.integers(12, 120, 2))
        col = int(rng.integers(20, 236))
        if rng.random() < 0.6:
            ang = float(rng.uniform(0, 180))
            box = cv2.boxPoints(((x, y), (w, h), ang)).astype(np.int32)
            cv2.fillConvexPoly(img, box, col, lineType=cv2.LINE_AA)
        else:
            x2, y2 = x + int(rng.integers(-300, 300)), y + int(rng.integers(-300, 300))
            cv2.line(img, (x, y), (x2, y2), col, int(rng.integers(2, 7)), lineType=cv2.LINE_AA)
    return img


def smooth_dem(size: int, seed: int, max_elev: int = 30) -> np.ndarray:
    """Smooth uint8 relief 0..max_elev (raw DEM units, no scaling: _shared.py:100-102)."""
    rng = np.random.default_rng(seed + 7919)
    n = size // 128 + 2
    coarse = rng.random((n, n)).astype(np.float32)
    up = cv2.resize(coarse, (n * 128, n * 128), interpolation=cv2.INTER_CUBIC)[:size, :size]
    up = (up - up.min()) / max(1e-6, float(up.max() - up.min()))
    return np.clip(np.round(up * max_elev), 0, 255).astype(np.uint8)


def rot_xyz(roll: float, pitch: float, yaw: float) -> np.ndarray:
    """World(raster ESD) -> camera rotation: yaw about z, then pitch about x, roll about y."""
    cz, sz = np.cos(yaw), np.sin(yaw)
    cx, sx = np.cos(pitch), np.sin(pitch)
    cy, sy = np.cos(roll), np.sin(roll)
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1.0]])
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    return ry @ rx @ rz


def tile_affine(tile_x0: float, tile_y0: float, gsd_m: float = 1.0) -> np.ndarray:
    """3x4 pixel->(lon,lat,alt) matrix the reference encodes as ``+proj=affine``
    (gis_node.py:618-634: z scale is negative because the raster frame is ESD)."""
    dlat = gsd_m / 110574.0
    dlon = gsd_m / (111320.0 * np.cos(np.radians(KSQL_LAT)))
    lon0 = KSQL_LON + tile_x0 * dlon
    lat0 = KSQL_LAT - tile_y0 * dlat
    return np.array(
        [[dlon, 0.0, 0.0, lon0], [0.0, -dlat, 0.0, lat0], [0.0, 0.0, -gsd_m, 0.0]], np.float64
    )


@dataclass
class SynthPair:
    frame: np.ndarray  # uint8 [H, W]  query camera image
    tile: np.ndarray  # uint8 [T, T]  orthophoto raster
    dem: np.ndarray  # uint8 [T, T]
    k: np.ndarray  # float64 [3, 3]
    r_gt: np.ndarray  # float64 [3, 3]  raster frame -> camera
    t_gt: np.ndarray  # float64 [3, 1]
    affine: np.ndarray  # float64 [3, 4]
    seed: int


def make_pair(
    ground: np.ndarray,
    seed: int,
    frame_hw: Tuple[int, int] = (720, 1280),
    tile_size: int = 1024,
    footprint_frac: float = 0.9,
    noise_sigma: float = 2.0,
    max_yaw_deg: float = 22.5,
    max_tilt_deg: float = 5.0,
) -> SynthPair:
    """One seeded pair: tile crop of ``ground`` + the frame a camera above it would see."""
    rng = np.random.default_rng(1_000_003 * (seed + 1))
    gsz = ground.shape[0]
    h, w = frame_hw
    margin = tile_size // 2 + 64
    tx0 = int(rng.integers(margin, gsz - tile_size - margin))
    ty0 = int(rng.integers(margin, gsz - tile_size - margin))
    tile = np.ascontiguousarray(ground[ty0 : ty0 + tile_size, tx0 : tx0 + tile_size])

    f = 0.32 * w
    k = np.array([[f, 0, w / 2.0], [0, f, h / 2.0], [0, 0, 1.0]])
    height = footprint_frac * tile_size * f / w  # footprint width = W/f * height
    cx = tile_size * (0.25 + 0.5 * rng.random())
    cy = tile_size * (0.25 + 0.5 * rng.random())
    yaw = np.radians(rng.uniform(-max_yaw_deg, max_yaw_deg))
    pitch = np.radians(rng.uniform(-max_tilt_deg, max_tilt_deg))
    roll = np.radians(rng.uniform(-max_tilt_deg, max_tilt_deg))
    r = rot_xyz(roll, pitch, yaw)
    c = np.array([[cx], [cy], [-height]])
    t = -r @ c

    # plane z=0 in raster coords -> image: H = K [r1 r2 t]; shift to ground-texture coordinates
    hmat = k @ np.column_stack((r[:, 0], r[:, 1], t[:, 0]))
    shift = np.array([[1, 0, -tx0], [0, 1, -ty0], [0, 0, 1.0]])
    frame = cv2.warpPerspective(
        ground, hmat @ shift, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT
    )
    if noise_sigma > 0:
        noise = rng.standard_normal((h, w)).astype(np.float32) * noise_sigma
        frame = np.clip(frame.astype(np.float32) + noise, 0, 255).astype(np.uint8)
    dem = np.zeros((tile_size, tile_size), np.uint8)
    return SynthPair(frame, tile, dem, k, r, t, tile_affine(tx0, ty0), seed)


def pair_stream(
    n: int,
    frame_hw: Tuple[int, int] = (720, 1280),
    tile_size: int = 1024,
    ground_size: int = 4096,
    ground_seed: int = 0,
    first_seed: int = 0,
    **kw,
) -> Iterator[SynthPair]:
    ground = ground_texture(ground_size, ground_seed)
    for i in range(n):
        yield make_pair(ground, first_seed + i, frame_hw, tile_size, **kw)


def synth_correspondences(
    seed: int,
    n_points: int = 500,
    outlier_frac: float = 0.25,
    noise_px: float = 0.5,
    tile_size: int = 1024,
    frame_hw: Tuple[int, int] = (720, 1280),
    relief: bool = True,
    outlier_min_px: float = 0.0,
) -> Dict[str, np.ndarray]:
    """Direct 2D-3D correspondences for the PnP stage (no images): reference keypoints on the
    raster, DEM heights, their projections into a camera with known pose + noise + outliers."""
    rng = np.random.default_rng(7_000_003 * (seed + 1))
    h, w = frame_hw
    f = 0.32 * w
    k = np.array([[f, 0, w / 2.0], [0, f, h / 2.0], [0, 0, 1.0]])
    dem = smooth_dem(tile_size, seed) if relief else np.zeros((tile_size, tile_size), np.uint8)
    height = 0.9 * tile_size * f / w
    c = np.array([[tile_size * rng.uniform(0.4, 0.6)], [tile_size * rng.uniform(0.4, 0.6)], [-height]])
    r = rot_xyz(*np.radians([rng.uniform(-5, 5), rng.uniform(-5, 5), rng.uniform(-22.5, 22.5)]))
    t = -r @ c
    pts_ref, pts_qry = [], []
    tries = 0
    while len(pts_ref) < n_points and tries < 100 * n_points:
        tries += 1
        xy = rng.uniform(4, tile_size - 5, 2).astype(np.float32)
        z = float(dem[int(np.floor(xy[1])), int(np.floor(xy[0]))])
        pc = r @ np.array([[xy[0]], [xy[1]], [z]], np.float64) + t
        if pc[2, 0] <= 1e-3:
            continue
        uv = (k @ pc)[:2, 0] / pc[2, 0]
        if not (4 <= uv[0] < w - 4 and 4 <= uv[1] < h - 4):
            continue
        pts_ref.append(xy)
        pts_qry.append(uv)
    ref = np.asarray(pts_ref, np.float32).reshape(-1, 2)
    qry = np.asarray(pts_qry, np.float64).reshape(-1, 2)
    qry += rng.standard_normal(qry.shape) * noise_px
    n_out = int(round(outlier_frac * len(ref)))
    if n_out:
        idx = rng.choice(len(ref), n_out, replace=False)
        true_uv = qry[idx].copy()
        new_uv = np.column_stack((rng.uniform(0, w, n_out), rng.uniform(0, h, n_out)))
        for _ in range(64):  # optionally keep outliers away from the consensus (unambiguous inlier set)
            close = np.linalg.norm(new_uv - true_uv, axis=1) < outlier_min_px
            if not close.any():
                break
            new_uv[close] = np.column_stack((rng.uniform(0, w, close.sum()), rng.uniform(0, h, close.sum())))
        qry[idx] = new_uv
    return dict(
        mkp_ref=ref, mkp_qry=qry.astype(np.float32), dem=dem, k=k, r_gt=r, t_gt=t,
        affine=tile_affine(1000.0, 1500.0),
    )

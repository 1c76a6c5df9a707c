"""PoseEstimator / compute_pose — drop-in for the PnP stage and the fused device path.

``compute_pose(camera_info, mkp_qry, mkp_ref, elevation) -> Optional[(r, t)]`` has the signature of
ros/gisnav/gisnav/core/_shared.py:89-125 (shared by PoseNode, pose_node.py:305, and TwistNode,
twist_node.py:289).  ``camera_info.k`` is float64[9] row-major (_shared.py:122); a plain 3x3 array
is accepted too.  "Cannot compute" returns ``None`` as the reference's callers expect
(pose_node.py:299-307).  ``estimate_from_images`` / ``estimate_batch`` run the whole of
pose_node.py:226-381 on the device for B independent (frame, raster) pairs.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from . import _lib
from .context import Context, ptr


@dataclass
class PoseResult:
    status: int
    r: np.ndarray  # [3,3] raster frame -> camera
    t: np.ndarray  # [3,1]
    ecef: np.ndarray  # [3] camera centre, metres (pose.pose.position, pose_node.py:362-365)
    quat: np.ndarray  # [4] x,y,z,w camera orientation in ECEF (pose_node.py:380-381)
    lla: np.ndarray  # [3] lon, lat, alt
    n_kp_qry: int
    n_kp_ref: int
    n_matches: int
    n_inliers: int
    best_hypothesis: int

    @property
    def ok(self) -> bool:
        return self.status == _lib.GNB_OK

    @property
    def camera_center(self) -> np.ndarray:
        """-R^T t, the camera position in the raster frame (pose_node.py:333-334)."""
        return (-self.r.T @ self.t).reshape(3)


def _k9(camera_info) -> np.ndarray:
    k = getattr(camera_info, "k", camera_info)
    k = np.ascontiguousarray(np.asarray(k, np.float64).reshape(9))
    return k


def _from_c(r: _lib.GnbPoseResult) -> PoseResult:
    return PoseResult(
        status=int(r.status), r=np.array(r.r, np.float64).reshape(3, 3), t=np.array(r.t, np.float64).reshape(3, 1),
        ecef=np.array(r.ecef, np.float64), quat=np.array(r.quat, np.float64), lla=np.array(r.lla, np.float64),
        n_kp_qry=int(r.n_kp_qry), n_kp_ref=int(r.n_kp_ref), n_matches=int(r.n_matches), n_inliers=int(r.n_inliers),
        best_hypothesis=int(r.best_hypothesis))


class PoseEstimator:
    def __init__(self, ctx: Optional[Context] = None, **ctx_kwargs):
        self.ctx = ctx or Context(**ctx_kwargs)

    # ---- the reference's compute_pose ---------------------------------------------------------
    def estimate(self, camera_info, mkp_qry: np.ndarray, mkp_ref: np.ndarray, elevation: Optional[np.ndarray],
                 return_inliers: bool = False):
        q = np.ascontiguousarray(mkp_qry, np.float32).reshape(-1, 2)
        r_ = np.ascontiguousarray(mkp_ref, np.float32).reshape(-1, 2)
        if q.shape[0] != r_.shape[0]:
            raise ValueError("mkp_qry and mkp_ref must have the same length")
        n = q.shape[0]
        k = _k9(camera_info)
        dem = None
        dh = dw = 0
        if elevation is not None:
            dem = np.ascontiguousarray(elevation, np.uint8)
            if dem.ndim == 3:
                dem = np.ascontiguousarray(dem[:, :, 0])
            dh, dw = dem.shape
        rm = np.zeros(9, np.float64)
        tv = np.zeros(3, np.float64)
        mask = np.zeros(max(n, 1), np.uint8)
        ninl = C.c_int(0)
        rc = self.ctx.check(self.ctx._lib.gnb_solve_pnp(self.ctx.handle, ptr(q), ptr(r_), n, ptr(dem), dh, dw, ptr(k), 0,
                                                        ptr(rm), ptr(tv), ptr(mask), C.byref(ninl)))
        if rc != _lib.GNB_OK:
            return None
        out = (rm.reshape(3, 3), tv.reshape(3, 1))
        if return_inliers:
            return out + (mask[:n].astype(bool),)
        return out

    def tail(self, r: np.ndarray, t: np.ndarray, affine: np.ndarray, ref_shape: Tuple[int, int]):
        """pose_node.py:333-381: (r, t) + pixel->WGS84 affine -> (ecef [3], quat xyzw [4], lla [3]) or None."""
        r9 = np.ascontiguousarray(r, np.float64).reshape(9)
        t3 = np.ascontiguousarray(t, np.float64).reshape(3)
        a12 = np.ascontiguousarray(np.asarray(affine, np.float64)[:3, :4]).reshape(12)
        ecef, quat, lla = np.zeros(3), np.zeros(4), np.zeros(3)
        rc = self.ctx.check(self.ctx._lib.gnb_geodetic_tail(self.ctx.handle, ptr(r9), ptr(t3), ptr(a12), int(ref_shape[0]),
                                                            int(ref_shape[1]), ptr(ecef), ptr(quat), ptr(lla)))
        if rc != _lib.GNB_OK:
            return None
        return ecef, quat, lla

    # ---- fused device path --------------------------------------------------------------------
    def estimate_batch(self, frames: np.ndarray, tiles: np.ndarray, dems: Optional[np.ndarray], ks: np.ndarray,
                       affines: np.ndarray) -> List[PoseResult]:
        """frames u8 [B,Hq,Wq], tiles u8 [B,Ht,Wt], dems u8 [B,Ht,Wt] or None, ks f64 [B,3,3],
        affines f64 [B,3,4] -> one PoseResult per pair (check ``.ok``)."""
        frames = np.ascontiguousarray(frames, np.uint8)
        tiles = np.ascontiguousarray(tiles, np.uint8)
        b, hq, wq = frames.shape
        bt, ht, wt = tiles.shape
        if b != bt:
            raise ValueError("frames and tiles must have the same batch size")
        ks = np.ascontiguousarray(np.asarray(ks, np.float64).reshape(b, 9))
        affines = np.ascontiguousarray(np.asarray(affines, np.float64).reshape(b, 12))
        if dems is not None:
            dems = np.ascontiguousarray(dems, np.uint8).reshape(b, ht, wt)
        res = (_lib.GnbPoseResult * b)()
        self.ctx.check(self.ctx._lib.gnb_pose_batch(self.ctx.handle, b, ptr(frames), hq, wq, ptr(tiles), ht, wt, ptr(dems),
                                                    ptr(ks), ptr(affines), 0, res))
        return [_from_c(x) for x in res]

    def estimate_batch_device(self, frames, tiles, dems, ks, affines) -> List[PoseResult]:
        """Same as :meth:`estimate_batch` with inputs already resident on the device as torch CUDA
        tensors (uint8 / float64); only the B result structs cross PCIe."""
        import torch

        b, hq, wq = frames.shape
        _, ht, wt = tiles.shape
        for t_ in (frames, tiles, ks, affines):
            assert t_.is_cuda and t_.is_contiguous()
        assert frames.dtype == torch.uint8 and tiles.dtype == torch.uint8
        assert ks.dtype == torch.float64 and affines.dtype == torch.float64
        torch.cuda.current_stream(frames.device).synchronize()
        res = (_lib.GnbPoseResult * b)()
        dptr = C.c_void_p(dems.data_ptr()) if dems is not None else None
        self.ctx.check(self.ctx._lib.gnb_pose_batch(self.ctx.handle, b, C.c_void_p(frames.data_ptr()), hq, wq,
                                                    C.c_void_p(tiles.data_ptr()), ht, wt, dptr, C.c_void_p(ks.data_ptr()),
                                                    C.c_void_p(affines.data_ptr()), 1, res))
        return [_from_c(x) for x in res]

    def estimate_candidates(self, query: np.ndarray, tiles: np.ndarray, tile_ids, dems: Optional[np.ndarray], camera_info,
                            affines: np.ndarray):
        """Candidate search (BASELINE config 4): one query frame against n reference rasters whose
        features are cached on the device by ``tile_ids`` (the reference caches the raster's features
        per stamp, pose_node.py:226-241).  ``tiles``: u8 [n,ht,wt], or a LIST of n [ht,wt] arrays / views — then only the
        rasters the cache does not hold are copied (a flyover re-uses its neighbours: no 8 MB gather per frame).
        -> (index of the best candidate or None, results, cache hits)."""
        query = np.ascontiguousarray(query, np.uint8)
        hq, wq = query.shape
        k = _k9(camera_info)
        lazy = isinstance(tiles, (list, tuple))
        n = len(tiles)
        ht, wt = tiles[0].shape
        affines = np.ascontiguousarray(np.asarray(affines, np.float64).reshape(n, 12))
        ids = None if tile_ids is None else np.ascontiguousarray(np.asarray(tile_ids, np.int64).reshape(n))
        if dems is not None:
            dems = np.ascontiguousarray(dems, np.uint8).reshape(n, ht, wt)
        res = (_lib.GnbPoseResult * n)()
        hits = C.c_int(0)
        if lazy:
            # a list of (possibly strided) views: only the rasters the device cache does not hold are made contiguous
            hit = np.zeros(n, np.int32)
            if ids is not None:
                self.ctx.check(self.ctx._lib.gnb_cache_lookup(self.ctx.handle, ptr(ids), n, ht, wt, ptr(hit)))
            keep = [None if hit[i] else np.ascontiguousarray(tiles[i], np.uint8) for i in range(n)]
            ptrs = (C.c_void_p * n)(*[None if t is None else t.ctypes.data for t in keep])
            self.ctx.check(self.ctx._lib.gnb_pose_candidates_ptrs(self.ctx.handle, ptr(query), hq, wq, n, C.cast(ptrs, C.c_void_p), ht, wt,
                                                                  ptr(ids), ptr(dems), ptr(k), ptr(affines), res, C.byref(hits)))
        else:
            tiles = np.ascontiguousarray(tiles, np.uint8)
            self.ctx.check(self.ctx._lib.gnb_pose_candidates(self.ctx.handle, ptr(query), hq, wq, n, ptr(tiles), ht, wt, ptr(ids),
                                                             ptr(dems), ptr(k), ptr(affines), res, C.byref(hits)))
        out = [_from_c(x) for x in res]
        ok = [i for i, r in enumerate(out) if r.ok]
        best = max(ok, key=lambda i: out[i].n_inliers) if ok else None
        return best, out, hits.value

    # ---- inspection of the last batch (accuracy reporting: keypoint / match sets per pair) ---------------
    def slot_keypoints(self, slot: int, with_descriptors: bool = False):
        """Keypoints (x, y) f32 [n,2] (+ descriptors f32 [n,256]) held in keypoint slot ``slot`` after a batch call:
        slots [0, max_batch) are the query frames, [max_batch, 2 max_batch) the rasters."""
        k = self.ctx.config.max_keypoints
        xy = np.empty((k, 2), np.float32)
        desc = np.empty((k, 256), np.float32) if with_descriptors else None
        n = C.c_int(0)
        self.ctx.check(self.ctx._lib.gnb_slot_keypoints(self.ctx.handle, slot, ptr(xy), None, ptr(desc), k, C.byref(n)))
        return (xy[: n.value], desc[: n.value]) if with_descriptors else xy[: n.value]

    def pair_matches(self, pair: int) -> np.ndarray:
        """Match index pairs int32 [n,2] (query keypoint, raster keypoint) of pair ``pair`` of the last batch call."""
        k = self.ctx.config.max_keypoints
        idx = np.empty((k, 2), np.int32)
        n = C.c_int(0)
        self.ctx.check(self.ctx._lib.gnb_pair_matches(self.ctx.handle, pair, ptr(idx), k, C.byref(n)))
        return idx[: n.value]

    def estimate_from_records(self, records, reference: np.ndarray, dem: Optional[np.ndarray], camera_info, affine: np.ndarray,
                              query_hw: Optional[Tuple[int, int]] = None) -> Optional[PoseResult]:
        """Query side as PRE-EXTRACTED keypoints in the packed wire format (``PointCloud2.data`` of
        ``OrthoStereoImage.query_sift``: one ``keypoint_dtype(256)`` record per keypoint, 1044 B; _shared.py:26-35,
        pose_node.py:207-213).  The bytes go to the device as they are and are unpacked there."""
        from .keypoint_record import keypoint_dtype

        rec = np.frombuffer(records, dtype=np.uint8) if isinstance(records, (bytes, bytearray, memoryview)) else np.ascontiguousarray(records).view(np.uint8).reshape(-1)
        step = keypoint_dtype(_lib.DESC_DIM).itemsize
        if rec.size % step:
            raise ValueError(f"record buffer of {rec.size} bytes is not a multiple of point_step {step}")
        reference = np.ascontiguousarray(reference, np.uint8)
        ht, wt = reference.shape
        if dem is not None:
            dem = np.ascontiguousarray(dem, np.uint8).reshape(ht, wt)
        k = _k9(camera_info)
        a12 = np.ascontiguousarray(np.asarray(affine, np.float64)[:3, :4]).reshape(12)
        hq, wq = (int(query_hw[0]), int(query_hw[1])) if query_hw is not None else (0, 0)
        res = _lib.GnbPoseResult()
        self.ctx.check(self.ctx._lib.gnb_pose_from_records(self.ctx.handle, ptr(rec), rec.size // step, step, _lib.DESC_DIM, hq, wq,
                                                           ptr(reference), ht, wt, ptr(dem), ptr(k), ptr(a12), C.byref(res)))
        out = _from_c(res)
        return out if out.ok else None

    def estimate_from_message(self, query, reference: np.ndarray, dem: Optional[np.ndarray], camera_info, crs: str,
                              query_hw: Optional[Tuple[int, int]] = None) -> Optional[PoseResult]:
        """Fields of an ``OrthoStereoImage`` message (ros/gisnav_msgs/msg/OrthoStereoImage.msg:14-18) as PoseNode
        receives them: mono8 reference / dem arrays, the ``+proj=affine`` CRS string (``_transformations.py:274-327``)
        and the query either as a mono8 image (2-D uint8 array) or as the ``query_sift`` PointCloud2 payload
        (bytes / 1-D array of packed 1044-byte keypoint records, pose_node.py:207-213)."""
        from .crs import proj_to_affine

        affine = proj_to_affine(crs)
        if isinstance(query, np.ndarray) and query.ndim == 2 and query.dtype == np.uint8:
            return self.estimate_from_images(query, reference, dem, camera_info, affine)
        return self.estimate_from_records(query, reference, dem, camera_info, affine, query_hw)

    def estimate_from_images(self, query: np.ndarray, reference: np.ndarray, dem: Optional[np.ndarray], camera_info,
                             affine: np.ndarray) -> Optional[PoseResult]:
        res = self.estimate_batch(query[None], reference[None], None if dem is None else dem[None],
                                  _k9(camera_info)[None], np.asarray(affine, np.float64)[None, :3, :4])[0]
        return res if res.ok else None


_default: Optional[PoseEstimator] = None


def compute_pose(camera_info, mkp_qry: np.ndarray, mkp_ref: np.ndarray, elevation: Optional[np.ndarray]
                 ) -> Optional[Tuple[np.ndarray, np.ndarray]]:
    """Module-level drop-in for ``gisnav.core._shared.compute_pose`` (_shared.py:89-125)."""
    global _default
    if _default is None:
        _default = PoseEstimator()
    return _default.estimate(camera_info, mkp_qry, mkp_ref, elevation)

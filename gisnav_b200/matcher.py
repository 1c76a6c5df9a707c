"""KeypointMatcher — drop-in for ``self._matcher`` in PoseNode.

Reference call site (ros/gisnav/gisnav/core/pose_node.py:285-287)::

    dists, match_indices = self._matcher(descs_qry, descs_ref, lafs_qry, lafs_ref)

with ``self._matcher = LightGlueMatcher("sift", {... "filter_threshold": 0.5 ...}).to(device).eval()``
(pose_node.py:109-121).  Returns ``(dists [K,1] float32, idxs [K,2] int64)``; column 0 indexes the
first descriptor set, column 1 the second (pose_node.py:296-297).  torch tensors are the
interchange type only; CUDA tensors are consumed and produced in place (no host round trip).
With no layer blob loaded the matcher is the LightGlue assignment head and the LAF arguments are
ignored.  After ``ctx.set_matcher_layers(blob)`` the transformer layers of the reference matcher run
in front of the head; the keypoint centres are then read from the LAFs
(``get_laf_center``: ``lafs[0, :, :, 2]``, pose_node.py:267-276) and normalised by ``hw1`` / ``hw2`` or,
when the caller passes none — the reference does not (pose_node.py:285-287) — by the largest keypoint
coordinate per axis, as kornia's LightGlueMatcher does.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from .context import Context, ptr


class KeypointMatcher:
    def __init__(self, ctx: Optional[Context] = None, **ctx_kwargs):
        self.ctx = ctx or Context(**ctx_kwargs)

    # kornia modules are chained as ``.to(device).eval()`` (pose_node.py:103-105); keep that working
    def to(self, *_args, **_kwargs):
        return self

    def eval(self):
        return self

    @staticmethod
    def _hw(kp: np.ndarray, hw):
        if hw is not None:
            return float(hw[0]), float(hw[1])
        if kp.shape[0] == 0:
            return 1.0, 1.0
        m = kp.max(axis=0)
        return float(m[1]), float(m[0])

    def match_arrays(self, desc1: np.ndarray, desc2: np.ndarray, kp1: Optional[np.ndarray] = None,
                     kp2: Optional[np.ndarray] = None, hw1=None, hw2=None):
        """desc f32 [N,256] / [M,256] (+ pixel keypoints f32 [N,2] / [M,2] when layers are loaded)."""
        d1 = np.ascontiguousarray(desc1, np.float32)
        d2 = np.ascontiguousarray(desc2, np.float32)
        if d1.ndim != 2 or d2.ndim != 2 or (d1.size and d1.shape[1] != _lib.DESC_DIM) or (d2.size and d2.shape[1] != _lib.DESC_DIM):
            raise ValueError("descriptors must be [N,256]")
        cap = max(1, min(d1.shape[0], d2.shape[0]))
        idx = np.empty((cap, 2), np.int64)
        sc = np.empty((cap,), np.float32)
        n = C.c_int(0)
        if kp1 is None or kp2 is None:
            if self.ctx.matcher_layers:
                raise ValueError("transformer layers are loaded: the matcher needs keypoints (pass lafs or kp1/kp2)")
            self.ctx.check(self.ctx._lib.gnb_match(self.ctx.handle, ptr(d1), d1.shape[0], ptr(d2), d2.shape[0], 0, ptr(idx),
                                                   ptr(sc), cap, C.byref(n)))
        else:
            k1 = np.ascontiguousarray(kp1, np.float32).reshape(-1, 2)
            k2 = np.ascontiguousarray(kp2, np.float32).reshape(-1, 2)
            if k1.shape[0] != d1.shape[0] or k2.shape[0] != d2.shape[0]:
                raise ValueError("one keypoint per descriptor expected")
            (h1, w1), (h2, w2) = self._hw(k1, hw1), self._hw(k2, hw2)
            self.ctx.check(self.ctx._lib.gnb_match_lightglue(
                self.ctx.handle, ptr(d1), ptr(k1), d1.shape[0], h1, w1, ptr(d2), ptr(k2), d2.shape[0], h2, w2, 0,
                ptr(idx), ptr(sc), cap, C.byref(n)))
        return sc[: n.value].reshape(-1, 1).copy(), idx[: n.value].copy()

    def refined_descriptors(self, side: int, n: int) -> np.ndarray:
        """Descriptors of side 0 / 1 after the transformer layers of the last call (parity hook)."""
        out = np.empty((n, _lib.DESC_DIM), np.float32)
        self.ctx.check(self.ctx._lib.gnb_refined_descriptors(self.ctx.handle, side, ptr(out), n))
        return out

    def __call__(self, desc1, desc2, lafs1=None, lafs2=None, hw1=None, hw2=None):
        import torch

        def centers(lafs):
            # kornia.feature.get_laf_center: lafs [1,N,2,3] -> [N,2]
            if lafs is None:
                return None
            a = lafs.detach().cpu().numpy() if isinstance(lafs, torch.Tensor) else np.asarray(lafs)
            return np.ascontiguousarray(a.reshape(-1, 2, 3)[:, :, 2], np.float32)

        if not isinstance(desc1, torch.Tensor):
            dists, idx = self.match_arrays(np.asarray(desc1), np.asarray(desc2), centers(lafs1), centers(lafs2), hw1, hw2)
            return torch.from_numpy(dists), torch.from_numpy(idx)
        if self.ctx.matcher_layers and desc1.is_cuda:
            # transformer layers loaded: the keypoint centres come from the LAFs, everything stays on the device
            if lafs1 is None or lafs2 is None:
                raise ValueError("transformer layers are loaded: the matcher needs lafs (keypoint centres)")
            d1 = desc1.detach().to(torch.float32).contiguous()
            d2 = desc2.detach().to(torch.float32).contiguous()
            k1 = lafs1.detach().to(device=d1.device, dtype=torch.float32).reshape(-1, 2, 3)[:, :, 2].contiguous()
            k2 = lafs2.detach().to(device=d1.device, dtype=torch.float32).reshape(-1, 2, 3)[:, :, 2].contiguous()

            def hw(kp, given):
                if given is not None:
                    return float(given[0]), float(given[1])
                if kp.shape[0] == 0:
                    return 1.0, 1.0
                m = kp.max(dim=0).values.tolist()      # two floats cross PCIe, like kornia's image-size inference
                return float(m[1]), float(m[0])

            (h1, w1), (h2, w2) = hw(k1, hw1), hw(k2, hw2)
            n1, n2 = d1.shape[0], d2.shape[0]
            cap = max(1, min(n1, n2))
            idx = torch.empty((cap, 2), dtype=torch.int64, device=d1.device)
            sc = torch.empty((cap,), dtype=torch.float32, device=d1.device)
            torch.cuda.current_stream(d1.device).synchronize()
            n = C.c_int(0)
            self.ctx.check(self.ctx._lib.gnb_match_lightglue(
                self.ctx.handle, C.c_void_p(d1.data_ptr()), C.c_void_p(k1.data_ptr()), n1, h1, w1, C.c_void_p(d2.data_ptr()),
                C.c_void_p(k2.data_ptr()), n2, h2, w2, 1, C.c_void_p(idx.data_ptr()), C.c_void_p(sc.data_ptr()), cap, C.byref(n)))
            return sc[: n.value].reshape(-1, 1), idx[: n.value]
        if self.ctx.matcher_layers:
            dists, idx = self.match_arrays(desc1.detach().cpu().numpy(), desc2.detach().cpu().numpy(), centers(lafs1),
                                           centers(lafs2), hw1, hw2)
            return torch.from_numpy(dists), torch.from_numpy(idx)
        if desc1.is_cuda:
            d1 = desc1.detach().to(torch.float32).contiguous()
            d2 = desc2.detach().to(torch.float32).contiguous()
            n1, n2 = d1.shape[0], d2.shape[0]
            cap = max(1, min(n1, n2))
            idx = torch.empty((cap, 2), dtype=torch.int64, device=d1.device)
            sc = torch.empty((cap,), dtype=torch.float32, device=d1.device)
            torch.cuda.current_stream(d1.device).synchronize()  # inputs were produced on torch's stream
            n = C.c_int(0)
            self.ctx.check(self.ctx._lib.gnb_match(self.ctx.handle, C.c_void_p(d1.data_ptr()), n1, C.c_void_p(d2.data_ptr()),
                                                   n2, 1, C.c_void_p(idx.data_ptr()), C.c_void_p(sc.data_ptr()), cap,
                                                   C.byref(n)))
            return sc[: n.value].reshape(-1, 1), idx[: n.value]
        dists, idx = self.match_arrays(desc1.detach().cpu().numpy(), desc2.detach().cpu().numpy())
        return torch.from_numpy(dists), torch.from_numpy(idx)


class BruteForceRatioMatcher:
    """Drop-in for TwistNode's matcher: ``self._bf = cv2.BFMatcher()`` ... ``knnMatch(desc_qry,
    desc_ref, k=2)`` followed by the ratio test ``m.distance < CONFIDENCE_THRESHOLD * n.distance``
    (ros/gisnav/gisnav/core/twist_node.py:54,95,248,263-267).  Returns the surviving (queryIdx,
    trainIdx) pairs in query order and ``m.distance``."""

    def __init__(self, ctx: Optional[Context] = None, ratio: float = 0.7, **ctx_kwargs):
        self.ctx = ctx or Context(**ctx_kwargs)
        self.ratio = ratio

    def knn_ratio_match(self, desc_qry: np.ndarray, desc_ref: np.ndarray):
        dq = np.ascontiguousarray(desc_qry, np.float32)
        dr = np.ascontiguousarray(desc_ref, np.float32)
        if dq.ndim != 2 or dr.ndim != 2 or (dq.size and dr.size and dq.shape[1] != dr.shape[1]):
            raise ValueError("descriptors must be [N,D] with equal D")
        dim = dq.shape[1] if dq.size else (dr.shape[1] if dr.size else 128)
        cap = max(1, dq.shape[0])
        idx = np.empty((cap, 2), np.int64)
        dist = np.empty((cap,), np.float32)
        n = C.c_int(0)
        self.ctx.check(self.ctx._lib.gnb_knn_ratio_match(self.ctx.handle, ptr(dq), dq.shape[0], ptr(dr), dr.shape[0], dim,
                                                         C.c_double(self.ratio), 0, ptr(idx), ptr(dist), cap, C.byref(n)))
        return idx[: n.value].copy(), dist[: n.value].copy()

    def knn_ratio_match_device(self, desc_qry, desc_ref):
        """Same with torch CUDA tensors in and out (no host round trip)."""
        import torch

        dq = desc_qry.detach().to(torch.float32).contiguous()
        dr = desc_ref.detach().to(torch.float32).contiguous()
        cap = max(1, dq.shape[0])
        idx = torch.empty((cap, 2), dtype=torch.int64, device=dq.device)
        dist = torch.empty((cap,), dtype=torch.float32, device=dq.device)
        torch.cuda.current_stream(dq.device).synchronize()
        n = C.c_int(0)
        self.ctx.check(self.ctx._lib.gnb_knn_ratio_match(self.ctx.handle, C.c_void_p(dq.data_ptr()), dq.shape[0], C.c_void_p(dr.data_ptr()),
                                                         dr.shape[0], dq.shape[1], C.c_double(self.ratio), 1, C.c_void_p(idx.data_ptr()),
                                                         C.c_void_p(dist.data_ptr()), cap, C.byref(n)))
        return idx[: n.value], dist[: n.value]

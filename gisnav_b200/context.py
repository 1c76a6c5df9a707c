"""Context = one gnb_ctx (device workspace + repacked weights) bound to one CUDA device."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, fields
from typing import Optional

import numpy as np

from . import _lib, weights as _weights


@dataclass
class Config:
    """Python mirror of ``struct gnb_config``.  Defaults follow the reference's class constants
    (ros/gisnav/gisnav/core/pose_node.py:60-72) where one exists."""

    max_keypoints: int = 1024  # MAX_KEYPOINTS, pose_node.py:66
    nms_radius: int = 4
    keypoint_threshold: float = 0.005
    border: int = 4
    match_threshold: float = 0.5  # CONFIDENCE_THRESHOLD, pose_node.py:60
    min_matches: int = 15  # MIN_MATCHES, pose_node.py:63
    ransac_iters: int = 2048  # reference: iterationsCount=10 + early exit (_shared.py:115)
    reproj_px: float = 8.0
    ransac_seed: int = 0
    refine: int = 1
    max_batch: int = 8
    max_image_h: int = 1088
    max_image_w: int = 1280
    conv_impl: Optional[int] = None  # 0 tcgen05 (product), 1 SIMT validation kernel; None = library default
    match_impl: Optional[int] = None
    tile_cache: int = 32  # reference-raster feature cache entries
    precision: int = 0  # 0 = bf16 operands (fast); 1 = split-bf16 x3 + fp32 heads/matcher (fp32-faithful)

    def __post_init__(self):
        if self.conv_impl is None or self.match_impl is None:
            d = _lib.GnbConfig()
            _lib.load().gnb_default_config(C.byref(d))
            if self.conv_impl is None:
                self.conv_impl = int(d.conv_impl)
            if self.match_impl is None:
                self.match_impl = int(d.match_impl)

    def to_c(self) -> _lib.GnbConfig:
        c = _lib.GnbConfig()
        for f in fields(self):
            setattr(c, f.name, getattr(self, f.name))
        return c


class Context:
    def __init__(self, config: Optional[Config] = None, weights: Optional[bytes] = None, device: int = 0,
                 weights_device_ptr: Optional[int] = None, weights_nbytes: int = 0):
        self._lib = _lib.load()
        self.config = config or Config()
        self.device = device
        self._h = C.c_void_p()
        cfg = self.config.to_c()
        if weights_device_ptr is not None:
            rc = self._lib.gnb_create(C.byref(cfg), C.c_void_p(weights_device_ptr), weights_nbytes, 1, device, C.byref(self._h))
        else:
            blob = weights if weights is not None else _weights.load()
            buf = (C.c_char * len(blob)).from_buffer_copy(blob)
            rc = self._lib.gnb_create(C.byref(cfg), C.cast(buf, C.c_void_p), len(blob), 0, device, C.byref(self._h))
        if rc != 0:
            raise _lib.GnbError(rc, self._lib.gnb_last_error(None).decode())

    @property
    def handle(self) -> C.c_void_p:
        if not self._h:
            raise RuntimeError("context destroyed")
        return self._h

    def check(self, rc: int) -> int:
        """Negative status => raise (usage/CUDA error); soft failures (>0) are returned."""
        if rc < 0:
            msg = self._lib.gnb_last_error(self._h).decode()
            if rc == _lib.GNB_E_RANGE:
                raise IndexError(msg)  # numpy raises IndexError at _shared.py:100-101
            raise _lib.GnbError(rc, msg)
        return rc

    def set_matcher_layers(self, blob: Optional[bytes]) -> None:
        """Load (or with ``None`` unload) the LightGlue transformer layers that run in front of the assignment
        head (``gisnav_b200.weights.pack_layers``; reference: ``n_layers=9``, pose_node.py:109-121)."""
        if blob is None:
            self.check(self._lib.gnb_set_matcher_layers(self._h, None, 0))
            return
        buf = (C.c_char * len(blob)).from_buffer_copy(blob)
        self.check(self._lib.gnb_set_matcher_layers(self._h, C.cast(buf, C.c_void_p), len(blob)))

    @property
    def matcher_layers(self) -> int:
        return int(self._lib.gnb_matcher_layers(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.gnb_launch_count(self._h))

    @property
    def stream_ptr(self) -> int:
        return int(self._lib.gnb_stream(self._h) or 0)

    def profile(self, on: bool) -> None:
        """Bracket every kernel launch with CUDA events on the library stream (bench roofline)."""
        self.check(self._lib.gnb_profile_enable(self._h, 1 if on else 0))

    def profile_read(self):
        """-> {kernel name: (total_ms, launches)} since the last read."""
        cap = 64
        names = C.create_string_buffer(64 * cap)
        ms = (C.c_float * cap)()
        cnt = (C.c_int64 * cap)()
        n = C.c_int(0)
        self.check(self._lib.gnb_profile_read(self._h, C.cast(names, C.c_void_p), C.cast(ms, C.c_void_p),
                                              C.cast(cnt, C.c_void_p), cap, C.byref(n)))
        out = {}
        for i in range(n.value):
            out[names.raw[64 * i: 64 * i + 64].split(b"\0")[0].decode()] = (float(ms[i]), int(cnt[i]))
        return out

    def close(self) -> None:
        if self._h:
            self._lib.gnb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ptr(a: Optional[np.ndarray]) -> Optional[C.c_void_p]:
    return None if a is None else C.c_void_p(a.ctypes.data)

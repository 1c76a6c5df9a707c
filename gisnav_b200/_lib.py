"""ctypes binding of libgisnav_b200.so (include/gisnav_b200.h).  No CPU fallback: if the shared
library is missing or no sm_100 device is present, the product path raises."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgisnav_b200.so")

GNB_OK = 0
GNB_SOFT_TOO_FEW_MATCHES = 1
GNB_SOFT_PNP_FAILED = 2
GNB_SOFT_OUT_OF_BOUNDS = 3
GNB_E_INVALID = -1
GNB_E_CUDA = -2
GNB_E_CAPACITY = -3
GNB_E_RANGE = -4
GNB_E_NO_DEVICE = -5
DESC_DIM = 256


class GnbConfig(C.Structure):
    """struct gnb_config (include/gisnav_b200.h)."""

    _fields_ = [
        ("max_keypoints", C.c_int32),
        ("nms_radius", C.c_int32),
        ("keypoint_threshold", C.c_float),
        ("border", C.c_int32),
        ("match_threshold", C.c_float),
        ("min_matches", C.c_int32),
        ("ransac_iters", C.c_int32),
        ("reproj_px", C.c_float),
        ("ransac_seed", C.c_uint32),
        ("refine", C.c_int32),
        ("max_batch", C.c_int32),
        ("max_image_h", C.c_int32),
        ("max_image_w", C.c_int32),
        ("conv_impl", C.c_int32),
        ("match_impl", C.c_int32),
        ("tile_cache", C.c_int32),
        ("precision", C.c_int32),
    ]


class GnbPoseResult(C.Structure):
    """struct gnb_pose_result (include/gisnav_b200.h)."""

    _fields_ = [
        ("status", C.c_int32),
        ("n_kp_qry", C.c_int32),
        ("n_kp_ref", C.c_int32),
        ("n_matches", C.c_int32),
        ("n_inliers", C.c_int32),
        ("best_hypothesis", C.c_int32),
        ("r", C.c_double * 9),
        ("t", C.c_double * 3),
        ("ecef", C.c_double * 3),
        ("quat", C.c_double * 4),
        ("lla", C.c_double * 3),
    ]


# every symbol include/gisnav_b200.h declares: name -> (restype, argtypes)
_VP, _I, _F = C.c_void_p, C.c_int, C.c_float
SIGNATURES = {
    "gnb_default_config": (_I, [C.POINTER(GnbConfig)]),
    "gnb_create": (_I, [C.POINTER(GnbConfig), _VP, C.c_size_t, _I, _I, C.POINTER(_VP)]),
    "gnb_destroy": (None, [_VP]),
    "gnb_last_error": (C.c_char_p, [_VP]),
    "gnb_get_config": (_I, [_VP, C.POINTER(GnbConfig)]),
    "gnb_launch_count": (C.c_int64, [_VP]),
    "gnb_stream": (_VP, [_VP]),
    "gnb_profile_enable": (_I, [_VP, _I]),
    "gnb_profile_read": (_I, [_VP, _VP, _VP, _VP, _I, C.POINTER(_I)]),
    "gnb_extract": (_I, [_VP, _VP, _I, _I, _I, _I, _VP, _VP, _VP, _I, C.POINTER(_I)]),
    "gnb_match": (_I, [_VP, _VP, _I, _VP, _I, _I, _VP, _VP, _I, C.POINTER(_I)]),
    "gnb_set_matcher_layers": (_I, [_VP, _VP, C.c_size_t]),
    "gnb_matcher_layers": (_I, [_VP]),
    "gnb_match_lightglue": (_I, [_VP, _VP, _VP, _I, _F, _F, _VP, _VP, _I, _F, _F, _I, _VP, _VP, _I, C.POINTER(_I)]),
    "gnb_refined_descriptors": (_I, [_VP, _I, _VP, _I]),
    "gnb_knn_ratio_match": (_I, [_VP, _VP, _I, _VP, _I, _I, C.c_double, _I, _VP, _VP, _I, C.POINTER(_I)]),
    "gnb_solve_pnp": (_I, [_VP, _VP, _VP, _I, _VP, _I, _I, _VP, _I, _VP, _VP, _VP, C.POINTER(_I)]),
    "gnb_geodetic_tail": (_I, [_VP, _VP, _VP, _VP, _I, _I, _VP, _VP, _VP]),
    "gnb_pose_batch": (_I, [_VP, _I, _VP, _I, _I, _VP, _I, _I, _VP, _VP, _VP, _I, C.POINTER(GnbPoseResult)]),
    "gnb_pose_from_records": (_I, [_VP, _VP, _I, _I, _I, _I, _I, _VP, _I, _I, _VP, _VP, _VP, C.POINTER(GnbPoseResult)]),
    "gnb_pose_candidates": (_I, [_VP, _VP, _I, _I, _I, _VP, _I, _I, _VP, _VP, _VP, _VP, C.POINTER(GnbPoseResult), C.POINTER(_I)]),
    "gnb_pose_candidates_ptrs": (_I, [_VP, _VP, _I, _I, _I, _VP, _I, _I, _VP, _VP, _VP, _VP, C.POINTER(GnbPoseResult), C.POINTER(_I)]),
    "gnb_cache_lookup": (_I, [_VP, _VP, _I, _I, _I, _VP]),
    "gnb_cache_clear": (_I, [_VP]),
    "gnb_rotate_crop": (_I, [_VP, _VP, _I, _VP, _I, _I, C.c_double, _I, _I, _I, _VP, _VP, _VP, _VP]),
    "gnb_dense": (_I, [_VP, _VP, _I, _I, _I, _VP, _VP]),
    "gnb_layer_activation": (_I, [_VP, C.c_char_p, _VP, C.c_size_t]),
    "gnb_dense_batch": (_I, [_VP, _VP, _I, _I, _I, _I, _I]),
    "gnb_layer_activation_at": (_I, [_VP, C.c_char_p, _I, _VP, C.c_size_t]),
    "gnb_slot_keypoints": (_I, [_VP, _I, _VP, _VP, _VP, _I, C.POINTER(_I)]),
    "gnb_pair_matches": (_I, [_VP, _I, _VP, _I, C.POINTER(_I)]),
    "gnb_select_keypoints": (_I, [_VP, _VP, _I, _I, _VP, _VP, _I, C.POINTER(_I)]),
    "gnb_sample_descriptors": (_I, [_VP, _VP, _I, _I, _VP, _I, _I, _I, _VP]),
    "gnb_match_scores": (_I, [_VP, _VP, _I, _VP, _I, _VP]),
    "gnb_ransac_debug": (_I, [_VP, _VP, _VP, C.POINTER(_I)]),
}

_lib: Optional[C.CDLL] = None


class GnbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libgisnav_b200 error {code}: {msg}")
        self.code = code


def load() -> C.CDLL:
    """dlopen the library and bind every declared symbol; raises if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C gisnav_b200/csrc`. gisnav_b200 has no CPU fallback."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib

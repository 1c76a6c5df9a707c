"""gisnav_b200 — B200 (sm_100a) implementation of GISNav's pose-estimation hot path.

Host-side mirror of the three call sites PoseNode uses:
``KeypointExtractor.detectAndCompute``, ``KeypointMatcher.__call__`` and ``compute_pose`` /
``PoseEstimator``, all backed by hand-written CUDA kernels in ``libgisnav_b200.so`` through the
C ABI in ``include/gisnav_b200.h``.  There is no CPU fallback.
"""
from .context import Config, Context  # noqa: F401
from .extractor import KeypointExtractor  # noqa: F401
from .keypoint_record import KEYPOINT_DTYPE, KEYPOINT_DTYPE_256  # noqa: F401
from .matcher import BruteForceRatioMatcher, KeypointMatcher  # noqa: F401
from .pose import PoseEstimator, PoseResult, compute_pose  # noqa: F401
from .stereo import StereoAligner  # noqa: F401
from .stream import FrameStream  # noqa: F401

__version__ = "0.1.0"

"""Several frames in flight on one GPU (BASELINE config 4: a flyover stream with a candidate search per frame).

One call of the hot path on ONE frame leaves most of a B200 idle in most of its kernels (top-K: one CTA; NMS passes,
descriptor head, 1/8-resolution layers: a fraction of a wave) and spends part of its time on the host (making the
missed rasters contiguous, ctypes marshalling, reading results).  Frames of a stream are independent, and so are
``Context`` objects (own streams, workspace, tensor maps, raster-feature cache: ``tests/test_gpu_parity.py::
test_two_live_contexts_do_not_share_state``), so ``FrameStream`` keeps ``n_contexts`` frames in flight: a pool of host
threads, one context each (ctypes releases the GIL during a call).  Results come back in submission order.

The reference node processes one message at a time under mutually exclusive callbacks
(``ros/gisnav/gisnav/__init__.py:140-154``); a replay, a multi-camera rig or a candidate search over many frames is
where this applies.  Measured (fp32-faithful mode, 8 candidates per frame, one GPU): 943 / 1 376 / 1 581 / 1 645
frames/s with 1 / 2 / 3 / 4 contexts.
"""
from __future__ import annotations

import queue
import threading
from concurrent.futures import Future
from typing import Callable, List, Optional, Sequence

from .context import Config, Context
from .pose import PoseEstimator


class FrameStream:
    """``submit(fn)`` runs ``fn(pose_estimator)`` on the next free context and returns a Future; ``map_frames`` is the
    ordered convenience form.  Close with ``close()`` (or use as a context manager)."""

    def __init__(self, n_contexts: int = 4, config: Optional[Config] = None, device: int = 0, weights: Optional[bytes] = None,
                 weights_device_ptr: Optional[int] = None, weights_nbytes: int = 0):
        if n_contexts < 1:
            raise ValueError("n_contexts must be >= 1")
        self.contexts: List[Context] = [Context(config, weights=weights, device=device, weights_device_ptr=weights_device_ptr,
                                                weights_nbytes=weights_nbytes) for _ in range(n_contexts)]
        self.estimators: List[PoseEstimator] = [PoseEstimator(c) for c in self.contexts]
        self._jobs: "queue.Queue" = queue.Queue()
        self._threads = [threading.Thread(target=self._worker, args=(pe,), daemon=True) for pe in self.estimators]
        for t in self._threads:
            t.start()

    def _worker(self, pe: PoseEstimator) -> None:
        while True:
            job = self._jobs.get()
            if job is None:
                return
            fn, fut = job
            if not fut.set_running_or_notify_cancel():
                continue
            try:
                fut.set_result(fn(pe))
            except BaseException as e:  # noqa: BLE001  (handed to the caller through the future)
                fut.set_exception(e)

    def submit(self, fn: Callable[[PoseEstimator], object]) -> Future:
        fut: Future = Future()
        self._jobs.put((fn, fut))
        return fut

    def map_frames(self, fns: Sequence[Callable[[PoseEstimator], object]]) -> list:
        """Run the callables (one per frame) with up to ``n_contexts`` in flight; results in the order given."""
        return [f.result() for f in [self.submit(fn) for fn in fns]]

    def for_each_context(self, fn: Callable[[PoseEstimator], object]) -> list:
        """Run ``fn`` once on EVERY context, serially (warm-up, loading matcher layers, clearing caches)."""
        return [fn(pe) for pe in self.estimators]

    @property
    def launch_count(self) -> int:
        return sum(c.launch_count for c in self.contexts)

    def close(self) -> None:
        for _ in self._threads:
            self._jobs.put(None)
        for t in self._threads:
            t.join()
        self._threads = []
        for c in self.contexts:
            c.close()
        self.contexts = []

    def __enter__(self) -> "FrameStream":
        return self

    def __exit__(self, *exc) -> None:
        self.close()

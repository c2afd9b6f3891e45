"""Frame-filter plug-ins with the reference's names and signatures
(utils/frame_filter.py of the reference): ``apply_roi``, ``downsample``, ``MotionFilter``.

Host arrays in, host arrays out -- exactly what ``StreamWorker._process_packet`` expects -- with
the pixel work done by the CUDA kernels.  CUDA tensors are accepted too and then stay on the
device, which is what the batched driver uses.  ROI masks are rasterised once per
(polygons, frame shape) and cached; the reference re-rasterises every frame.
"""

from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _native
from .runtime import FrameStager, get_handle
from .types import MotionFilterConfig

_mask_cache: Dict[tuple, object] = {}
_stagers: Dict[int, FrameStager] = {}


def _stager(h) -> FrameStager:
    s = _stagers.get(id(h))
    if s is None:
        s = _stagers[id(h)] = FrameStager(h)
    return s


def roi_mask(polygons, height: int, width: int, handle: Optional[_native.Handle] = None):
    """CUDA uint8 [H,W] mask: union of ``cv2.fillPoly(mask, [poly], 255)`` (frame_filter.py:46-49)."""
    h = handle if handle is not None else get_handle()
    key = (id(h), height, width, tuple(tuple((int(x), int(y)) for x, y in poly) for poly in polygons))
    m = _mask_cache.get(key)
    if m is None:
        m = h.roi_rasterize(polygons, height, width)
        if len(_mask_cache) > 256:
            _mask_cache.clear()
        _mask_cache[key] = m
    return m


def _is_tensor(x) -> bool:
    return hasattr(x, "is_cuda")


def apply_roi(frame, polygons, handle: Optional[_native.Handle] = None):
    """frame_filter.py:43-50."""
    if not polygons:
        return frame
    h = handle if handle is not None else get_handle()
    dev = frame if _is_tensor(frame) else _stager(h).upload([frame])[0]
    out = h.apply_mask(dev, roi_mask(polygons, dev.shape[0], dev.shape[1], h))
    return out if _is_tensor(frame) else out.cpu().numpy()


def downsample(frame, scale: float, handle: Optional[_native.Handle] = None):
    """frame_filter.py:53-57."""
    if scale >= 0.999:
        return frame
    h = handle if handle is not None else get_handle()
    fh, fw = frame.shape[:2]
    dev = frame if _is_tensor(frame) else _stager(h).upload([frame])[0]
    out = h.resize([dev], [(int(fh * scale), int(fw * scale))])[0]
    return out if _is_tensor(frame) else out.cpu().numpy()


class MotionFilter:
    """frame_filter.py:19-40.  ``previous_gray`` lives on the device (two ping-pong buffers)."""

    def __init__(self, config: MotionFilterConfig, frame_shape: Tuple[int, ...] = (0, 0, 0),
                 handle: Optional[_native.Handle] = None):
        self.config = config
        self.h = handle if handle is not None else get_handle()
        self._bufs: List = []
        self._cur = -1  # index of the buffer holding previous_gray; -1: no previous frame yet
        self.last_count: Optional[int] = None
        self.alpha = 1.0 / max(1, config.history)

    @property
    def previous_gray(self) -> Optional[np.ndarray]:
        return None if self._cur < 0 else self._bufs[self._cur].cpu().numpy()

    def _ensure(self, hgt: int, wid: int) -> None:
        t = self.h.torch
        if not self._bufs or tuple(self._bufs[0].shape) != (hgt, wid):
            if self._bufs and self._cur >= 0:
                # cv2.absdiff would raise on a size change; mirror that instead of guessing
                raise ValueError("MotionFilter: frame size changed")
            self._bufs = [t.empty((hgt, wid), dtype=t.uint8, device=self.h.device) for _ in range(2)]

    def buffers(self, hgt: int, wid: int):
        """(prev or None, next) device buffers for the batched kernel call."""
        self._ensure(hgt, wid)
        nxt = 0 if self._cur < 0 else 1 - self._cur
        return (None if self._cur < 0 else self._bufs[self._cur]), self._bufs[nxt]

    def commit(self, count: int) -> bool:
        """Advance the state after a kernel call that wrote ``next``; returns should_process."""
        first = self._cur < 0
        self._cur = 0 if first else 1 - self._cur
        if first:
            self.last_count = None
            return True
        self.last_count = int(count)
        hgt, wid = self._bufs[0].shape
        motion_ratio = float(count) / float(hgt * wid)
        return motion_ratio >= self.config.threshold

    def advance(self) -> None:
        """Device-side gates: the kernel wrote ``next``; the decision itself is taken on the device
        (``b200va_gates_decide``), so only the ping-pong index moves here."""
        self._cur = 0 if self._cur < 0 else 1 - self._cur
        self.last_count = None

    def should_process(self, frame, roi_mask=None) -> bool:
        if not self.config.enable:
            return True
        dev = frame if _is_tensor(frame) else _stager(self.h).upload([frame])[0]
        prev, nxt = self.buffers(dev.shape[0], dev.shape[1])
        counts = self.h.motion([dev], [prev], [nxt], [roi_mask] if roi_mask is not None else None)
        return self.commit(int(counts.cpu()[0]))

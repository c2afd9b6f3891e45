"""Process-wide plumbing: the default library handle and host<->device staging.

One process drives one GPU (one ``Handle`` per device).  PyTorch supplies device memory, pinned
host memory and the current CUDA stream; nothing here computes.
"""

from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _native

_handles: Dict[int, "_native.Handle"] = {}


def default_device() -> int:
    import torch

    if "LOCAL_RANK" in os.environ and torch.cuda.device_count() > 1:
        return int(os.environ["LOCAL_RANK"]) % torch.cuda.device_count()
    return torch.cuda.current_device() if torch.cuda.is_available() else 0


def get_handle(device: Optional[int] = None, **overrides) -> "_native.Handle":
    """The shared handle of a device (created on first use).  Capacities come from
    ``B200VA_MAX_*`` environment variables or keyword overrides on the first call."""
    dev = default_device() if device is None else int(device)
    h = _handles.get(dev)
    if h is None:
        cfg = dict(max_batch=int(os.environ.get("B200VA_MAX_BATCH", 64)),
                   max_anchors=int(os.environ.get("B200VA_MAX_ANCHORS", 25200)),
                   max_candidates=int(os.environ.get("B200VA_MAX_CANDIDATES", 4096)),
                   max_dets=int(os.environ.get("B200VA_MAX_DETS", 1024)),
                   max_streams=int(os.environ.get("B200VA_MAX_STREAMS", 64)),
                   max_tracks=int(os.environ.get("B200VA_MAX_TRACKS", 4096)))
        cfg.update(overrides)
        h = _native.Handle(device=dev, **cfg)
        _handles[dev] = h
    return h


def reset_handles() -> None:
    for h in _handles.values():
        h.close()
    _handles.clear()


class FrameStager:
    """Pinned host ring + device buffers for decoded frames (numpy HWC uint8 -> CUDA)."""

    def __init__(self, handle: "_native.Handle"):
        self.h = handle
        self._pinned: Dict[tuple, list] = {}
        self._turn: Dict[tuple, int] = {}

    def upload(self, frames: Sequence) -> List:
        """Return CUDA uint8 [H,W,3] tensors for ``frames``; CUDA tensors pass through untouched."""
        t = self.h.torch
        out = []
        slot_of: Dict[tuple, int] = {}
        for f in frames:
            if t.is_tensor(f):
                if not f.is_cuda:
                    f = f.to(self.h.device, non_blocking=True)
                out.append(f)
                continue
            a = np.asarray(f)
            if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
                raise ValueError("frames must be uint8 arrays of shape [H, W, 3]")
            if not a.flags["C_CONTIGUOUS"]:
                a = np.ascontiguousarray(a)
            key = a.shape
            k = slot_of.get(key, 0)
            slot_of[key] = k + 1
            ring = self._pinned.setdefault(key, [])
            # two generations per slot so a frame still in flight is never overwritten
            gen = self._turn.get(key, 0)
            idx = 2 * k + gen
            while len(ring) <= idx:
                ring.append(t.empty(key, dtype=t.uint8, pin_memory=True))
            pinned = ring[idx]
            pinned.numpy()[...] = a
            out.append(pinned.to(self.h.device, non_blocking=True))
        for key in slot_of:
            self._turn[key] = 1 - self._turn.get(key, 0)
        return out

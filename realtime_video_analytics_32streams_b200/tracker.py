"""Tracker plug-in: ``IouTracker``'s contract (tracker.py:36-126 of the reference) with the track
table resident in HBM and the matching done by ``b200va_tracker_update``.

``update(stream_name, detections)`` keeps the reference signature: Python ``Detection`` objects in
(their float64 fields go down unchanged), all surviving ``Track`` objects out, one id counter
shared by every stream.  ``update_batch`` is the additive batched entry: device SoA detections
straight from ``B200Detector.predict_batch_device`` for many streams in one launch.
"""

from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

from . import _native
from .runtime import get_handle
from .types import Track


class B200IouTracker:
    def __init__(self, config, handle: Optional[_native.Handle] = None):
        self.config = config
        self.h = handle if handle is not None else get_handle()
        self._slots: Dict[str, int] = {}
        self._out = None

    # ---- helpers ---------------------------------------------------------------------------
    def slot_of(self, stream_name: str) -> int:
        s = self._slots.get(stream_name)
        if s is None:
            s = len(self._slots)
            if s >= self.h.cfg.max_streams:
                raise _native.B200VAError(_native.ERR_CAPACITY, f"more than {self.h.cfg.max_streams} streams")
            self._slots[stream_name] = s
        return s

    def _cfg(self):
        c = self.config
        return int(c.max_age), int(c.min_hits), float(c.max_iou_distance)

    def reset(self, stream_name: str) -> None:
        if stream_name in self._slots:
            self.h.tracker_reset(self._slots[stream_name])

    @staticmethod
    def tracks_from_soa(out, index: int) -> List[Track]:
        n = int(out["count"][index])
        return [Track(int(out["track_id"][index, i]), int(out["cls"][index, i]), float(out["conf"][index, i]),
                      tuple(float(v) for v in out["bbox_xyxy"][index, i]), int(out["age"][index, i]),
                      int(out["hits"][index, i])) for i in range(n)]

    @staticmethod
    def tracks_from_soa_all(host) -> List[List[Track]]:
        """All rows of a host SoA -> lists of Track objects (vectorised conversions)."""
        counts = host["count"].tolist()
        ids, cls, conf = host["track_id"].tolist(), host["cls"].tolist(), host["conf"].tolist()
        box, age, hits = host["bbox_xyxy"].tolist(), host["age"].tolist(), host["hits"].tolist()
        return [[Track(ids[b][i], cls[b][i], conf[b][i], tuple(box[b][i]), age[b][i], hits[b][i]) for i in range(n)]
                for b, n in enumerate(counts)]

    @staticmethod
    def soa_to_host(out, kmax: Optional[int] = None):
        """Device SoA -> dict of numpy arrays (one synchronising copy per field)."""
        counts = out["count"].cpu().numpy()
        k = int(counts.max()) if len(counts) else 0
        if kmax is not None:
            k = min(k, kmax)
        host = {"count": counts, "new_count": out["new_count"].cpu().numpy()}
        for key in ("track_id", "cls", "conf", "bbox_xyxy", "age", "hits"):
            host[key] = out[key][:, :max(k, 1)].cpu().numpy()
        return host

    # ---- reference API ---------------------------------------------------------------------
    def update(self, stream_name: str, detections: Iterable) -> List[Track]:
        """tracker.py:50-95."""
        t = self.h.torch
        dets = list(detections)
        n = len(dets)
        cap = max(n, 1)
        box = np.zeros((1, cap, 4), dtype=np.float64)
        conf = np.zeros((1, cap), dtype=np.float64)
        cls = np.zeros((1, cap), dtype=np.int32)
        for i, d in enumerate(dets):
            box[0, i] = d.bbox_xyxy
            conf[0, i] = d.confidence
            cls[0, i] = d.class_id
        dev = self.h.device
        soa = {"bbox_xyxy": t.from_numpy(box).to(dev), "conf": t.from_numpy(conf).to(dev),
               "cls": t.from_numpy(cls).to(dev), "count": t.tensor([n], dtype=t.int32, device=dev)}
        max_age, min_hits, thr = self._cfg()
        out = self.h.tracker_update([self.slot_of(stream_name)], soa, max_age, min_hits, thr, f64=True, out=self._one())
        return self.tracks_from_soa(self.soa_to_host(out), 0)

    def _one(self):
        if self._out is None:
            self._out = self.h.alloc_tracks(1)
        return self._out

    # ---- batched API -----------------------------------------------------------------------
    def update_batch(self, stream_names: Sequence[str], dets, det_scale=None, skip=None, id_base=None, out=None):
        """One launch for many streams.  ``dets`` is the device SoA of ``postprocess`` (row i belongs to
        ``stream_names[i]``); ``skip[i]`` reproduces ``tracker.update(stream, [])``.  Ids are drawn
        from the shared counter in the order of ``stream_names``.  Returns the device SoA."""
        slots = [self.slot_of(s) for s in stream_names]
        max_age, min_hits, thr = self._cfg()
        return self.h.tracker_update(slots, dets, max_age, min_hits, thr, det_scale=det_scale, skip=skip,
                                     id_base=id_base, out=out)

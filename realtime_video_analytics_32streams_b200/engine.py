"""Batched per-tick driver: ``StreamWorker._process_packet`` (pipeline.py:143-262 of the
reference) for many streams at once.

One tick takes the latest decoded frame of every stream and runs, in the reference's order,
ROI -> downsample -> motion gate -> adaptive-FPS gate -> predict -> rescale -> float64 re-threshold
-> tracker -> adaptive state update.  The per-stream gates are O(streams) integers and stay on the
host exactly as the reference computes them; everything that touches pixels, head tensors or
track tables is one batched kernel launch per step.  Skipped streams still age their tracks
(``_skip_frame``: ``tracker.update(stream, [])``).
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from . import _native
from .detector import B200Detector
from .frame_filter import MotionFilter, roi_mask
from .runtime import FrameStager, get_handle
from .tracker import B200IouTracker
from .types import Detection, FrameResult, MotionFilterConfig, Track


@dataclass
class _StreamState:
    """Host-side gate state of one stream (pipeline.py:88-116)."""

    cfg: object
    motion: Optional[MotionFilter] = None
    frame_index: int = 0
    idle_frames: int = 0
    process_every: int = 1
    max_process_every: int = 1
    idle_tolerance: int = 0
    adaptive: bool = False

    def __post_init__(self):
        s = self.cfg
        self.adaptive = bool(getattr(s, "adaptive_fps", False))
        if self.adaptive:
            target = getattr(s, "target_fps", None) or 30.0
            min_fps = max(getattr(s, "min_target_fps", 5.0), 1.0)
            self.max_process_every = max(1, int(round(target / min_fps)))
            self.idle_tolerance = max(int(getattr(s, "idle_frame_tolerance", 60)), 1)

    def adjust(self, n_det: int, n_trk: int) -> None:
        """``_adjust_adaptive_state`` (pipeline.py:242-262)."""
        if not self.adaptive:
            return
        if n_det > 0 or n_trk > 0:
            self.idle_frames = 0
            self.process_every = 1
        else:
            self.idle_frames += 1
            if self.idle_frames >= self.idle_tolerance:
                self.process_every = max(self.max_process_every, 1)


class HotPathEngine:
    """pre + post + track for a set of streams, one tick at a time."""

    def __init__(self, streams: Sequence, detector_config, tracker_config, infer: Callable,
                 handle: Optional[_native.Handle] = None, input_hw=None, build_objects: bool = True):
        self.h = handle if handle is not None else get_handle()
        self.streams = list(streams)
        self.detector = B200Detector(detector_config, input_hw=input_hw, infer=infer, handle=self.h, fold_filter=True)
        self.tracker = B200IouTracker(tracker_config, handle=self.h)
        self.state: Dict[str, _StreamState] = {s.name: _StreamState(s) for s in self.streams}
        self.stager = FrameStager(self.h)
        self.build_objects = build_objects
        n = max(len(self.streams), 1)
        self._dets = self.h.alloc_dets(n)
        self._tracks = self.h.alloc_tracks(n)
        self._changed = self.h.torch.empty((n,), dtype=self.h.torch.int32, device=self.h.device)
        self.last_soa = None  # device SoA of the last tick (tracks), for callers that skip objects

    # --------------------------------------------------------------------------------------
    def tick(self, frames: Sequence, frame_ids: Optional[Sequence[int]] = None,
             infer_ctx=None) -> List[FrameResult]:
        """``frames[i]`` is the new frame of ``streams[i]`` (host array or CUDA tensor) or None
        when that stream delivered nothing this tick."""
        t = self.h.torch
        live = [i for i, f in enumerate(frames) if f is not None]
        if not live:
            return []
        dev_frames = self.stager.upload([frames[i] for i in live])
        names = [self.streams[i].name for i in live]
        states = [self.state[n] for n in names]
        for st in states:
            st.frame_index += 1
        ids = [frame_ids[i] if frame_ids is not None else self.state[self.streams[i].name].frame_index for i in live]

        # 1-2. ROI (fused into the consumers) and downsample
        masks, work, ratios = [], [], []
        for st, f in zip(states, dev_frames):
            polys = getattr(st.cfg, "roi_polygons", None) or []
            masks.append(roi_mask(polys, f.shape[0], f.shape[1], self.h) if polys else None)
            ratios.append(float(getattr(st.cfg, "downsample_ratio", 1.0)))
        down = [k for k, r in enumerate(ratios) if r < 0.999]
        work = list(dev_frames)
        work_masks = list(masks)
        if down:
            sizes = [(int(dev_frames[k].shape[0] * ratios[k]), int(dev_frames[k].shape[1] * ratios[k])) for k in down]
            small = self.h.resize([dev_frames[k] for k in down], sizes, [masks[k] for k in down])
            for k, s in zip(down, small):
                work[k] = s
                work_masks[k] = None  # the ROI is already baked into the downsampled frame

        # 3. motion gate (one launch for every stream that has it enabled, one count read-back)
        skip_reason: List[Optional[str]] = [None] * len(live)
        mot = [k for k, st in enumerate(states) if getattr(st.cfg, "motion_filter", False)]
        if mot:
            prevs, nexts = [], []
            for k in mot:
                st = states[k]
                if st.motion is None:
                    st.motion = MotionFilter(MotionFilterConfig(enable=True, threshold=st.cfg.motion_threshold),
                                             tuple(work[k].shape), handle=self.h)
                p, n = st.motion.buffers(work[k].shape[0], work[k].shape[1])
                prevs.append(p)
                nexts.append(n)
            counts = self.h.motion([work[k] for k in mot], prevs, nexts, [work_masks[k] for k in mot],
                                   self._changed[:len(mot)]).cpu().numpy()
            for j, k in enumerate(mot):
                if not states[k].motion.commit(int(counts[j])):
                    skip_reason[k] = "motion"

        # 4. adaptive-FPS gate (pipeline.py:165-170)
        for k, st in enumerate(states):
            if skip_reason[k] is None and st.adaptive and st.process_every > 1:
                if (st.frame_index - 1) % st.process_every != 0:
                    skip_reason[k] = "adaptive"

        # 5-8. predict on the active subset, then the tracker over every live stream
        active = [k for k in range(len(live)) if skip_reason[k] is None]
        skipped = [k for k in range(len(live)) if skip_reason[k] is not None]
        order = active + skipped  # skipped streams create no tracks, so id order is unaffected
        n_act = len(active)
        dets = {k_: v[:len(order)] for k_, v in self._dets.items()}
        if n_act:
            tensor, metas = self.h.preprocess([work[k] for k in active], self.detector.input_hw, self.detector._fmt,
                                              [work_masks[k] for k in active])
            head = self.detector._infer(tensor) if infer_ctx is None else self.detector._infer_fn(tensor, infer_ctx)
            head = self.detector._as_head(head)
            if head.dim() != 3 or head.shape[0] != n_act:
                raise ValueError(f"infer returned {tuple(head.shape)} for a batch of {n_act}")
            self.detector._run_post(head, metas, {k_: v[:n_act] for k_, v in self._dets.items()})
        scale = [1.0 / max(ratios[k], 1e-6) if ratios[k] < 0.999 else 1.0 for k in order]
        out = {k_: v[:len(order)] for k_, v in self._tracks.items()}
        self.tracker.update_batch([names[k] for k in order], dets,
                                  det_scale=scale if any(ratios[k] < 0.999 for k in order) else None,
                                  skip=[0] * n_act + [1] * len(skipped), out=out)
        self.last_soa = (order, dets, out)

        # 9-13. read back the counts the adaptive state machine needs; build objects on request
        host_tr = B200IouTracker.soa_to_host(out) if self.build_objects else \
            {"count": out["count"].cpu().numpy()}
        det_counts = dets["count"][:n_act].cpu().numpy() if n_act else np.zeros(0, np.int32)
        host_det = None
        if self.build_objects and n_act and int(det_counts.max()) > 0:
            kmax = int(det_counts.max())
            host_det = (dets["bbox_xyxy"][:n_act, :kmax].cpu().numpy(), dets["conf"][:n_act, :kmax].cpu().numpy(),
                        dets["cls"][:n_act, :kmax].cpu().numpy())
        results: List[Optional[FrameResult]] = [None] * len(live)
        for pos, k in enumerate(order):
            st = states[k]
            processed = pos < n_act
            n_det = int(det_counts[pos]) if processed else 0
            n_trk = int(host_tr["count"][pos])
            st.adjust(n_det, n_trk)
            res = FrameResult(names[k], ids[k], processed, skip_reason[k])
            if self.build_objects:
                if processed and host_det is not None:
                    sc = scale[pos]
                    box, conf, cls = host_det
                    res.detections = [Detection(names[k], ids[k], int(cls[pos, i]), float(conf[pos, i]),
                                                tuple(float(v) * sc if sc != 1.0 else float(v) for v in box[pos, i]))
                                      for i in range(n_det)]
                res.tracks = B200IouTracker.tracks_from_soa(host_tr, pos)
            else:
                res.detections = n_det  # type: ignore[assignment]
                res.tracks = n_trk  # type: ignore[assignment]
            results[k] = res
        return [r for r in results if r is not None]

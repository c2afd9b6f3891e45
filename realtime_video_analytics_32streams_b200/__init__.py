"""B200-native pre + post + track hot path of skygazer42/realtime-video-analytics-32streams.

Public surface (names follow the reference so call sites read the same):

* ``B200Detector``   -- ``BaseDetector`` contract: ``predict(packet)``, plus ``predict_batch``.
* ``B200UltralyticsDetector`` -- the same contract with the pre / post semantics of ``UltralyticsDetector``.
* ``B200IouTracker`` -- ``IouTracker`` contract: ``update(stream_name, detections)``, plus ``update_batch``.
* ``apply_roi``, ``downsample``, ``MotionFilter``, ``MotionFilterConfig`` -- ``utils/frame_filter.py``.
* ``HotPathEngine``  -- the batched per-tick driver (``StreamWorker._process_packet`` for N streams).
* ``register_with_reference`` -- makes ``detector.backend: b200`` / ``tracker.type: b200_iou`` selectable
  in the reference's YAML config when the reference package is importable.

All compute lives in ``lib/libb200va.so`` (C ABI: ``include/b200va.h``); importing a compute symbol
fails loudly when the library is missing.  Nothing in here imports ``oracle/``.
"""

from .types import Detection, DetectorConfig, FramePacket, MotionFilterConfig, StreamConfig, Track, TrackerConfig

__all__ = ["Detection", "DetectorConfig", "FramePacket", "FrameResult", "MotionFilterConfig", "StreamConfig", "Track",
           "TrackerConfig", "B200Detector", "B200UltralyticsDetector", "B200IouTracker", "HotPathEngine", "MotionFilter", "apply_roi",
           "downsample", "filter_detections", "get_handle", "register_with_reference", "TickCollector", "B200KafkaSink"]

_LAZY = {
    "B200Detector": "detector", "B200UltralyticsDetector": "detector", "filter_detections": "detector", "B200IouTracker": "tracker",
    "HotPathEngine": "engine", "FrameResult": "engine", "MotionFilter": "frame_filter", "apply_roi": "frame_filter",
    "downsample": "frame_filter", "roi_mask": "frame_filter", "get_handle": "runtime",
    "register_with_reference": "integration", "TickCollector": "collector", "B200KafkaSink": "sinks",
}


def __getattr__(name):
    mod = _LAZY.get(name)
    if mod is None:
        raise AttributeError(name)
    import importlib

    return getattr(importlib.import_module(f"{__name__}.{mod}"), name)

"""Plain records and config holders with the reference's field names.

The product package cannot import the reference (it is not installed next to it), so the handful
of dataclasses the hot path touches are declared here with identical field names and defaults:
``Detection`` (detector.py:32-40), ``Track`` (tracker.py:18-33), ``FramePacket``
(video_stream.py:26-33) and the knobs of ``DetectorConfig`` / ``TrackerConfig`` / ``StreamConfig``
that the path reads (config.py:57-73, 129-142, 194-201).  Every class in this package only uses
attribute access, so the reference's own config / packet objects can be passed in unchanged.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, List, Optional, Tuple


@dataclass(slots=True)
class Detection:
    stream_name: str
    frame_id: int
    class_id: int
    confidence: float
    bbox_xyxy: Tuple[float, float, float, float]


@dataclass(slots=True)
class Track:
    track_id: int
    class_id: int
    confidence: float
    bbox_xyxy: Tuple[float, float, float, float]
    age: int = 0
    hits: int = 0
    # optional temporal fields the Kafka sink reads (kafka_sink.py:107-123); unused on this path
    action_label: Optional[str] = None
    temporal_score: Optional[float] = None
    sequence_start_frame: Optional[int] = None
    sequence_end_frame: Optional[int] = None


@dataclass(slots=True)
class StreamConfig:
    name: str
    url: str = ""
    enabled: bool = True
    target_fps: Optional[float] = None
    batch_size: int = 1
    detector_id: Optional[str] = None
    roi_polygons: Optional[List[List[Tuple[int, int]]]] = None
    motion_filter: bool = False
    motion_threshold: float = 0.02
    downsample_ratio: float = 1.0
    adaptive_fps: bool = False
    min_target_fps: float = 5.0
    idle_frame_tolerance: int = 60


@dataclass(slots=True)
class FramePacket:
    stream: Any  # StreamConfig-like: needs .name
    frame: Any  # numpy [H,W,3] uint8 BGR (host) or a CUDA uint8 tensor of the same layout
    frame_id: int
    timestamp: float = 0.0


@dataclass(slots=True)
class DetectorConfig:
    model_path: str = "yolov8n.pt"
    device: str = "auto"
    backend: str = "b200"
    model_type: str = "yolov8"
    confidence_threshold: float = 0.5
    iou_threshold: float = 0.45
    classes: Optional[List[int]] = None
    half: bool = False
    warmup: bool = True
    input_size: Optional[List[int]] = None  # H, W


@dataclass(slots=True)
class TrackerConfig:
    type: str = "b200_iou"
    max_age: int = 30
    max_iou_distance: float = 0.7
    min_hits: int = 3


@dataclass(slots=True)
class MotionFilterConfig:
    enable: bool = False
    history: int = 5
    threshold: float = 0.02
    blur_kernel: Tuple[int, int] = (5, 5)

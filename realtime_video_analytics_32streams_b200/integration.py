"""Registration shim for the reference package.

The reference selects its detector by ``detector.backend`` (whitelist: config.py:155-157, dispatch:
detector.py:54-96) and hard-codes ``IouTracker`` (pipeline.py:452).  ``register_with_reference``
patches those three places at import time so that a YAML with ``backend: b200`` (or
``b200_ultralytics``: the Ultralytics pre / post semantics of detector.py:106-179) and
``tracker.type: b200_iou`` runs this package's kernels behind the reference's own pipeline, and
swaps the frame-filter functions the pipeline imported by name (pipeline.py:33).  With ``batched=True`` it also
installs the tick collector (``collector.py``): the reference's per-stream workers then share ONE batched
``HotPathEngine.tick`` instead of calling the kernels 32 times at batch 1 (pipeline.py:118-212, 460-515).
INTEGRATION.md shows the equivalent two-line source change for maintainers who prefer a patch.
"""

from __future__ import annotations

from typing import Callable, Optional

_ORIGINALS = {}  # (module or class, attribute) -> the reference's own object, for unregister_from_reference


def _patch(owner, name, value):
    _ORIGINALS.setdefault((owner, name), getattr(owner, name))
    setattr(owner, name, value)


def unregister_from_reference() -> None:
    """Put back everything ``register_with_reference`` replaced (tests; a process that wants the stock pipeline again)."""
    for (owner, name), value in list(_ORIGINALS.items()):
        setattr(owner, name, value)
    _ORIGINALS.clear()
    try:
        import realtime_analytics.pipeline as rpipe

        rpipe.StreamWorker._b200va_batched = False
    except Exception:  # pragma: no cover
        pass


def register_with_reference(infer_factory: Optional[Callable] = None, batched: bool = False,
                            engine_factory: Optional[Callable] = None, max_wait_s: float = 0.010,
                            gpu_sink: bool = False) -> None:
    """``infer_factory(config) -> callable`` builds the detector forward for a DetectorConfig
    (e.g. loads the YOLO weights with PyTorch); it is required for ``backend: b200``.
    ``batched=True``: all stream workers of a pipeline share one tick (``collector.install``); ``engine_factory(streams,
    detector, pipeline_config)`` overrides the engine it drives (default: ``HotPathEngine`` on the detector's handle) and
    ``max_wait_s`` bounds how long a tick waits for a late stream.
    ``gpu_sink=True``: the pipeline's ``KafkaSink`` (pipeline.py:453) becomes ``B200KafkaSink`` -- same messages, the
    preview's downscale / boxes and the event serialisation done by the library (sinks.py)."""
    import realtime_analytics.config as rcfg  # the reference, must be importable
    import realtime_analytics.detector as rdet
    import realtime_analytics.pipeline as rpipe

    from .detector import B200Detector, B200UltralyticsDetector
    from .frame_filter import MotionFilter, apply_roi, downsample
    from .tracker import B200IouTracker

    # 1. whitelist the backend id (config.py:155-157 builds the set inside validate())
    orig_validate = rcfg.DetectorConfig.validate

    def validate(self):
        if self.backend in ("b200", "b200_ultralytics"):
            backend, self.backend = self.backend, "tensorrt"
            try:
                orig_validate(self)
            finally:
                self.backend = backend
        else:
            orig_validate(self)

    _patch(rcfg.DetectorConfig, "validate", validate)

    # 2. dispatch (detector.py:54-96)
    orig_create = rdet.create_detector

    def create_detector(config):
        if config.backend.lower() in ("b200", "b200_ultralytics"):
            if infer_factory is None:
                raise RuntimeError(f"backend '{config.backend}' needs register_with_reference(infer_factory=...)")
            cls = B200UltralyticsDetector if config.backend.lower() == "b200_ultralytics" else B200Detector
            return cls(config, infer=infer_factory(config))
        return orig_create(config)

    _patch(rdet, "create_detector", create_detector)
    _patch(rpipe, "create_detector", create_detector)

    # 3. tracker (pipeline.py:452) keyed on tracker.type, and the frame filters (pipeline.py:33)
    class _Tracker:
        def __new__(cls, config):
            if getattr(config, "type", "") == "b200_iou":
                return B200IouTracker(config)
            return rpipe._ReferenceIouTracker(config)

    if not hasattr(rpipe, "_ReferenceIouTracker"):
        rpipe._ReferenceIouTracker = rpipe.IouTracker
    _patch(rpipe, "IouTracker", _Tracker)
    _patch(rpipe, "apply_roi", apply_roi)
    _patch(rpipe, "downsample", downsample)
    _patch(rpipe, "MotionFilter", MotionFilter)

    # 4. one tick for all workers (pipeline.py:118-212, 460-515)
    if batched:
        from . import collector

        for owner, name in ((rpipe.AnalyticsPipeline, "__init__"), (rpipe.AnalyticsPipeline, "wait_closed"),
                            (rpipe.StreamWorker, "__init__"), (rpipe.StreamWorker, "run"), (rpipe.StreamWorker, "_process_packet")):
            _ORIGINALS.setdefault((owner, name), getattr(owner, name))
        collector.install(rpipe, engine_factory or collector.default_engine_factory, max_wait_s)

    # 5. egress (pipeline.py:453, sinks/kafka_sink.py)
    if gpu_sink:
        from .sinks import B200KafkaSink

        _patch(rpipe, "KafkaSink", B200KafkaSink)

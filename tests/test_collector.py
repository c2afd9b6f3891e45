"""The in-pipeline tick collector (collector.py).

CPU part (runs in the build container, where the reference tree is importable): the REAL reference
``StreamWorker`` objects (pipeline.py:88-262), patched by ``collector.install``, share one tick per frame period and
produce, per stream, exactly the side effects the unpatched workers produce one frame at a time --
``metrics.update_counters``, ``kafka.send_tracks`` (track ids, boxes, hits), ``health.update_success`` /
``update_error`` and the adaptive-FPS fields.  The engine behind the collector is an oracle-backed stand-in there
(no GPU); the GPU part runs the same collector over the real ``HotPathEngine``.
"""
import asyncio
import os
import sys
import types

import numpy as np
import pytest

import golden_util as G
from oracle import hotpath as O
from realtime_video_analytics_32streams_b200 import synth
from realtime_video_analytics_32streams_b200.collector import TickCollector, install, make_process_packet

H, W, IN_HW = 108, 192, (64, 64)
CONF, IOU = 0.35, 0.5
N_STREAMS, N_FRAMES = 4, 20
POLYS = [[(10, 8), (180, 12), (170, 100), (20, 95)]]


def _stream_kwargs(i):
    return dict(name=f"cam{i}", roi_polygons=POLYS if i % 2 == 0 else None, motion_filter=(i != 3),
                motion_threshold=0.02, downsample_ratio=1.0 if i != 1 else 0.75, adaptive_fps=True, target_fps=25,
                min_target_fps=5, idle_frame_tolerance=3)


def _scene(i):
    return synth.MotionScene(61 + i, H, W, rect=30, speed=11, static=(i == 2))


def _head(i, t):
    n_obj = 5 if (i != 2 and (t < 6 or t > 13)) else 0
    return synth.synth_head(7000 + 10 * t + i, 20, 256, n_obj, dup=2, input_hw=IN_HW)[None]


class Recorder:
    """Stub sinks: record every call a worker makes, per stream."""

    def __init__(self):
        self.calls = []

    def update_counters(self, **kw):
        self.calls.append(("metrics", kw["stream"], kw["frames_processed"], kw["detections_emitted"], kw["active_tracks"]))

    async def send_tracks(self, stream_name, frame_id, tracks, frame=None):
        arr = G.tracks_arrays(list(tracks))
        self.calls.append(("kafka", stream_name, frame_id, {k: v.copy() for k, v in arr.items()}, frame is not None))


class Health:
    def __init__(self, log, name):
        self.log, self.name = log, name

    def update_success(self, dt):
        assert dt >= 0.0
        self.log.append(("health_ok", self.name))

    def update_error(self):
        self.log.append(("health_err", self.name))


class OracleEngine:
    """Stand-in for HotPathEngine on a box without a GPU: same ``streams`` / ``tick`` contract, CPU oracle inside."""

    def __init__(self, streams, conf, iou, trk_cfg, head_fn):
        self.streams = list(streams)
        self.t = {s.name: 0 for s in self.streams}
        tracker = O.IouTracker(*trk_cfg)
        self.workers = {}
        for i, s in enumerate(self.streams):
            spec = O.StreamSpec(name=s.name, roi_polygons=s.roi_polygons, motion_filter=s.motion_filter,
                                motion_threshold=s.motion_threshold, downsample_ratio=s.downsample_ratio,
                                adaptive_fps=s.adaptive_fps, target_fps=s.target_fps, min_target_fps=s.min_target_fps,
                                idle_frame_tolerance=s.idle_frame_tolerance)
            self.workers[s.name] = O.StreamWorker(spec, (lambda tensor, idx, i=i, n=s.name: head_fn(i, self.t[n])), tracker,
                                                  conf, iou, None, IN_HW, False, backend="cv2")
        self.batches = []

    def tick(self, frames, frame_ids=None):
        out = []
        self.batches.append([f is not None for f in frames])
        for i, (s, f) in enumerate(zip(self.streams, frames)):
            if f is None:
                continue
            self.t[s.name] = frame_ids[i]
            w = self.workers[s.name]
            r = w.process(f)
            out.append(types.SimpleNamespace(stream_name=s.name, frame_id=frame_ids[i], processed=r.processed,
                                             skip_reason=r.skip_reason, n_detections=len(r.detections),
                                             n_tracks=len(r.tracks), tracks=r.tracks, detections=r.detections,
                                             adaptive_state=(w.process_every, w.idle_frames)))
        return out


def _same_calls(got, want):
    assert len(got) == len(want), (len(got), len(want))
    for g, w in zip(got, want):
        assert g[0] == w[0] and g[1] == w[1], (g[:2], w[:2])
        if g[0] == "metrics":
            assert g == w, (g, w)
        elif g[0] == "kafka":
            assert g[2] == w[2] and g[4] == w[4]
            for k in w[3]:
                assert np.array_equal(g[3][k], w[3][k]), (g[1], g[2], k)


def test_collector_drives_the_real_reference_workers_like_their_own_per_frame_path():
    ref = "/root/" + "reference/src"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present")
    sys.path.insert(0, ref)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    try:
        import realtime_analytics.config as rcfg
        import realtime_analytics.pipeline as rpipe
        from realtime_analytics.detector import _TensorRTBaseDetector
        from realtime_analytics.tracker import IouTracker
        from realtime_analytics.video_stream import FramePacket
    except Exception as exc:  # pragma: no cover
        pytest.skip(f"reference not importable: {exc}")

    class StubDetector(_TensorRTBaseDetector):  # the reference's numpy pre / post with the forward replaced by a lookup
        def __init__(self, config, input_hw):
            super().__init__(config, input_hw)
            self.head = None

        def _infer(self, tensor):
            return self.head

    det_cfg = rcfg.DetectorConfig(backend="tensorrt", confidence_threshold=CONF, iou_threshold=IOU)
    trk_cfg = rcfg.TrackerConfig(max_age=3, max_iou_distance=0.5, min_hits=1)
    streams = [rcfg.StreamConfig(url="x", **_stream_kwargs(i)) for i in range(N_STREAMS)]
    scenes = [_scene(i) for i in range(N_STREAMS)]
    frames = [[scenes[i].frame(t) for i in range(N_STREAMS)] for t in range(N_FRAMES)]

    def make_workers(tracker, detector, rec, log):
        ws = []
        for s in streams:
            ctx = rpipe.StreamWorkerContext(stream=s, detector=detector, tracker=tracker, kafka=rec, metrics=rec,
                                            health=Health(log, s.name))
            w = rpipe.StreamWorker(ctx)
            w._maybe_save_snapshot = lambda *a, **k: None
            ws.append(w)
        return ws

    # (A) the unmodified reference, one frame at a time, streams in order
    rec_a, log_a = Recorder(), []
    det_a = StubDetector(det_cfg, IN_HW)
    workers_a = make_workers(IouTracker(trk_cfg), det_a, rec_a, log_a)
    state_a = []

    async def drive_a():
        for t in range(N_FRAMES):
            for i, w in enumerate(workers_a):
                det_a.head = _head(i, t)
                await w._process_packet(FramePacket(streams[i], frames[t][i], t, 0.0))
            state_a.append([(w._process_every, w._idle_frames, w._frame_index) for w in workers_a])

    asyncio.run(drive_a())
    assert any(c[0] == "kafka" for c in rec_a.calls) and any(c[0] == "metrics" and c[3] == 0 for c in rec_a.calls)

    # (B) the same workers patched: every frame period is ONE engine tick
    saved = {k: getattr(rpipe.StreamWorker, k) for k in ("_process_packet", "run", "__init__")}
    saved_p = {k: getattr(rpipe.AnalyticsPipeline, k) for k in ("__init__", "wait_closed")}
    engines = []

    def engine_factory(strs, detector, cfg):
        e = OracleEngine(strs, CONF, IOU, (3, 0.5, 1), _head)
        engines.append(e)
        return e

    try:
        install(rpipe, engine_factory, max_wait_s=5.0)
        install(rpipe, engine_factory)  # idempotent
        cfg = rcfg.PipelineConfig(streams=streams, detector=det_cfg, tracker=trk_cfg, kafka=rcfg.KafkaSinkConfig(enabled=False))
        pipe = rpipe.AnalyticsPipeline(cfg)
        rec_b, log_b = Recorder(), []
        det_b = types.SimpleNamespace(config=det_cfg)
        workers_b = make_workers(pipe.tracker, det_b, rec_b, log_b)
        state_b = []

        async def drive_b():
            for t in range(N_FRAMES):
                await asyncio.gather(*[w._process_packet(FramePacket(streams[i], frames[t][i], t, 0.0))
                                       for i, w in enumerate(workers_b)])
                state_b.append([(w._process_every, w._idle_frames, w._frame_index) for w in workers_b])
            for col in pipe._b200va["collectors"].values():
                assert col.ticks == N_FRAMES and col.frames == N_FRAMES * N_STREAMS
                await col.close()

        asyncio.run(drive_b())
    finally:
        for k, v in saved.items():
            setattr(rpipe.StreamWorker, k, v)
        for k, v in saved_p.items():
            setattr(rpipe.AnalyticsPipeline, k, v)
        rpipe.StreamWorker._b200va_batched = False
    assert len(engines) == 1 and [s.name for s in engines[0].streams] == [s.name for s in streams]
    assert all(all(b) for b in engines[0].batches) and len(engines[0].batches) == N_FRAMES
    # same side effects per stream, in the same order (calls of different streams interleave differently: the batched
    # workers resume after the tick, so compare stream by stream)
    for s in streams:
        _same_calls([c for c in rec_b.calls if c[1] == s.name], [c for c in rec_a.calls if c[1] == s.name])
        assert [c for c in log_b if c[1] == s.name] == [c for c in log_a if c[1] == s.name]
    assert state_b == state_a


class _Packet:
    def __init__(self, stream, frame, frame_id):
        self.stream, self.frame, self.frame_id = stream, frame, frame_id


def test_collector_partial_ticks_errors_and_retired_streams():
    """A late stream does not hold a tick back longer than max_wait_s, an engine failure reaches every waiting worker
    (and is counted by its health tracker), and a retired stream is not waited for."""
    names = ["a", "b", "c"]
    streams = [types.SimpleNamespace(name=n) for n in names]

    class Engine:
        def __init__(self):
            self.streams, self.seen, self.fail = streams, [], False

        def tick(self, frames, ids):
            if self.fail:
                raise ValueError("boom")
            self.seen.append([f is not None for f in frames])
            return [types.SimpleNamespace(stream_name=s.name, processed=True, n_detections=1, n_tracks=2, tracks=[], frame_id=i,
                                          adaptive_state=None) for s, f, i in zip(streams, frames, ids) if f is not None]

    eng = Engine()
    col = TickCollector(eng, max_wait_s=0.05)
    log, rec = [], Recorder()
    proc = make_process_packet(lambda w: col)
    workers = [types.SimpleNamespace(ctx=types.SimpleNamespace(metrics=rec, kafka=rec, health=Health(log, n)), _frame_index=0,
                                     _maybe_save_snapshot=lambda *a, **k: None, _process_every=1, _idle_frames=0) for n in names]

    async def drive():
        # tick 1: only a and b deliver -> fires after max_wait_s without c
        await asyncio.gather(*[proc(workers[i], _Packet(streams[i], np.zeros((2, 2, 3), np.uint8), 1)) for i in (0, 1)])
        assert eng.seen == [[True, True, False]]
        # c retires: a tick of a and b is now "full" immediately
        col.retire("c")
        col.max_wait_s = 30.0
        await asyncio.wait_for(asyncio.gather(*[proc(workers[i], _Packet(streams[i], np.zeros((2, 2, 3), np.uint8), 2))
                                                for i in (0, 1)]), timeout=5.0)
        assert eng.seen[-1] == [True, True, False]
        # a second packet of the same stream before the tick is refused (one outstanding packet per worker)
        eng.fail = True
        res = await asyncio.gather(*[proc(workers[i], _Packet(streams[i], np.zeros((2, 2, 3), np.uint8), 3)) for i in (0, 1)],
                                   return_exceptions=True)
        assert all(isinstance(r, ValueError) for r in res)
        eng.fail = False
        with pytest.raises(KeyError):
            await col.process("nope", _Packet(streams[0], None, 0))
        await col.close()
        with pytest.raises(RuntimeError):
            await col.process("a", _Packet(streams[0], None, 0))

    asyncio.run(drive())
    assert [c for c in log if c[0] == "health_err"] == [("health_err", "a"), ("health_err", "b")]
    assert sum(1 for c in log if c[0] == "health_ok") == 4 and col.ticks == 2


@pytest.mark.gpu
def test_collector_over_the_gpu_engine_matches_the_oracle_pipeline():
    import torch
    from realtime_video_analytics_32streams_b200 import (DetectorConfig, HotPathEngine, StreamConfig, TrackerConfig,
                                                         _native)

    h = _native.Handle(device=0, max_batch=N_STREAMS, max_anchors=256, max_candidates=256, max_dets=64,
                       max_streams=N_STREAMS, max_tracks=128)
    try:
        streams = [StreamConfig(**_stream_kwargs(i)) for i in range(N_STREAMS)]
        tick = {"t": 0}
        eng = HotPathEngine(streams, DetectorConfig(confidence_threshold=CONF, iou_threshold=IOU),
                            TrackerConfig(max_age=3, max_iou_distance=0.5, min_hits=1),
                            infer=lambda tensor: torch.from_numpy(
                                np.concatenate([_head(i, tick["t"]) for i in eng.active_streams])).to(h.device),
                            handle=h, input_hw=IN_HW)
        ora = OracleEngine(streams, CONF, IOU, (3, 0.5, 1), _head)
        col = TickCollector(eng, max_wait_s=5.0)
        scenes = [_scene(i) for i in range(N_STREAMS)]

        async def drive():
            for t in range(N_FRAMES):
                tick["t"] = t
                frames = [scenes[i].frame(t) for i in range(N_STREAMS)]
                got = await asyncio.gather(*[col.process(s.name, _Packet(s, frames[i], t)) for i, s in enumerate(streams)])
                want = ora.tick(frames, [t] * N_STREAMS)
                for g, w in zip(got, want):
                    assert (g.stream_name, g.processed, g.skip_reason, g.n_detections, g.n_tracks, g.adaptive_state) == \
                           (w.stream_name, w.processed, w.skip_reason, w.n_detections, w.n_tracks, w.adaptive_state), (t, g.stream_name)
                    ga, wa = G.tracks_arrays(g.tracks), G.tracks_arrays(w.tracks)
                    for k in wa:
                        assert np.array_equal(ga[k], wa[k]), (t, g.stream_name, k)
            assert col.ticks == N_FRAMES
            await col.close()

        asyncio.run(drive())
        h.poll_status()
    finally:
        h.close()

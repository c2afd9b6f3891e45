"""Parity at the BASELINE.json shapes the small golden cases do not reach:

* config 4 END TO END: 32 streams x 2160x3840, per-stream hexagon + triangle ROI, motion gate, adaptive FPS --
  ``HotPathEngine`` against ``oracle.StreamWorker`` (pipeline.py:143-262 restated; OpenCV back end) stream by stream
  and tick by tick: processed flag, skip reason, process_every, idle_frames, detections, full track tables, ids.
* config 5 for the whole 200-frame run: dense heads (~1800 candidates -> ~300 kept per frame), long-lived tracks,
  ``max_tracks=4096``, two streams sharing the id counter.  The scene is ``DenseScene(orbit=3, jitter=0.7)``: the
  default linear drift walks the 300 objects into each other after ~40 frames (the oracle alone then holds 28 of the
  first frame's 319 tracks at tick 200 and > 3600 rows), which is a different workload from "long-lived tracks".
"""
import numpy as np
import pytest

import golden_util as G
from oracle import hotpath as O
from realtime_video_analytics_32streams_b200 import synth

pytestmark = pytest.mark.gpu


class FastMotionScene(synth.MotionScene):
    """``synth.MotionScene`` with the per-frame noise drawn from two cached fields (even / odd frames) instead of a
    fresh 25 M-sample draw per frame: consecutive frames still differ by noise in [-6, 6] everywhere (below the
    motion gate's 25) and by the moving rectangles, and a 4K frame costs a copy plus three rectangle fills."""

    _noise = {}

    def frame(self, t: int) -> np.ndarray:
        key = (self.h, self.w)
        if key not in FastMotionScene._noise:
            rng = np.random.default_rng(4242)
            FastMotionScene._noise[key] = [rng.integers(-3, 4, size=(self.h, self.w, 3), dtype=np.int16) for _ in range(2)]
        if not hasattr(self, "_base"):
            self._base = [np.clip(self.background.astype(np.int16) + n, 0, 255).astype(np.uint8)
                          for n in FastMotionScene._noise[key]]
        img = self._base[t & 1].copy()
        if not self.static:
            for k in range(len(self.pos)):
                x = int((self.pos[k, 0] + t * self.vel[k, 0]) % max(self.w - self.rect, 1))
                y = int((self.pos[k, 1] + t * self.vel[k, 1]) % max(self.h - self.rect, 1))
                img[y:y + self.rect, x:x + self.rect] = self.colors[k]
        return img


def _config4_head(s: int, t: int):
    """Objects come and go per stream so that the adaptive-FPS path is walked in both directions."""
    if s % 4 == 1:
        n_obj = 6 if (t < 6 or t >= 24) else 0  # long idle stretch -> process_every rises, then resets
    elif s % 4 == 2:
        n_obj = 0 if 10 <= t < 16 else 9
    else:
        n_obj = 12
    return synth.synth_head(40000 + 100 * t + s, 84, 8400, n_obj, dup=3)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("device_gates", [False, True])
def test_config4_engine_32x4k_roi_motion_adaptive_vs_oracle_stream_worker(device_gates):
    """``device_gates=True``: gates decided on the device (b200va_gates_decide / _commit) and ticks PIPELINED -- tick
    t + 1 is submitted before tick t is collected, which the host-side gates cannot allow for these streams."""
    import torch
    from realtime_video_analytics_32streams_b200 import (DetectorConfig, HotPathEngine, StreamConfig, TrackerConfig,
                                                         _native)

    S, T, HH, WW = 32, 30, 2160, 3840
    conf, iou = 0.35, 0.5
    trk_cfg = dict(max_age=3, max_iou_distance=0.5, min_hits=1)
    h = _native.Handle(device=0, max_batch=S, max_anchors=8400, max_candidates=2048, max_dets=512, max_streams=S,
                       max_tracks=1024)
    try:
        scenes, specs, streams = [], [], []
        for s in range(S):
            static = s % 5 == 4  # every 5th stream never moves: motion skip from its second frame on
            # rectangle size / speed / per-stream threshold chosen so that the changed-pixel ratio inside the ROI
            # straddles the gate: some streams always pass, some never, most flip from tick to tick
            scenes.append(FastMotionScene(500 + s, HH, WW, static=static, rect=700 + 60 * (s % 4), speed=90 + 70 * (s % 5)))
            kw = dict(name=f"cam-4k-{s:02d}", roi_polygons=synth.synth_polygons(600 + s, HH, WW), motion_filter=True,
                      motion_threshold=(0.02, 0.01, 0.004)[s % 3], downsample_ratio=1.0, adaptive_fps=True, target_fps=25, min_target_fps=5,
                      idle_frame_tolerance=3)
            specs.append(O.StreamSpec(**kw))
            streams.append(StreamConfig(**kw))
        tick = {"t": 0}
        ora_trk = O.IouTracker(trk_cfg["max_age"], trk_cfg["max_iou_distance"], trk_cfg["min_hits"])
        workers = [O.StreamWorker(specs[s], (lambda tensor, idx, s=s: _config4_head(s, tick["t"])[None]), ora_trk, conf, iou,
                                  None, (640, 640), False, backend="cv2") for s in range(S)]
        eng = HotPathEngine(streams, DetectorConfig(confidence_threshold=conf, iou_threshold=iou), TrackerConfig(**trk_cfg),
                            infer=lambda tensor: torch.from_numpy(
                                np.stack([_config4_head(s, tick["t"]) for s in eng.active_streams])).to(h.device),
                            handle=h, input_hw=(640, 640), device_gates=device_gates)
        seen = {"motion": 0, "adaptive": 0, "processed": 0, "raised": 0}
        pending = None
        for t in range(T + 1):
            if t < T:
                tick["t"] = t
                frames = [sc.frame(t) for sc in scenes]
                if device_gates:
                    ctx, ctx_frames = eng.submit(frames), frames
                else:
                    got, got_t, got_frames = eng.tick(frames), t, frames
            if device_gates:
                if pending is None:
                    pending = (ctx, t, ctx_frames)
                    continue
                got, got_t, got_frames = eng.collect(pending[0]), pending[1], pending[2]
                pending = (ctx, t, ctx_frames) if t < T else None
            elif t == T:
                break
            tick["t"] = got_t  # the oracle's infer callback reads it
            for s in range(S):
                want = workers[s].process(got_frames[s])
                r = got[s]
                assert (r.processed, r.skip_reason) == (want.processed, want.skip_reason), (got_t, s)
                assert r.adaptive_state == (workers[s].process_every, workers[s].idle_frames), (got_t, s)
                wc, wf, wb = G.dets_arrays(want.detections)
                gc, gf, gb = G.dets_arrays(r.detections)
                assert np.array_equal(gc, wc) and np.array_equal(gf, wf) and np.array_equal(gb, wb), (got_t, s, "detections")
                wt, gt = G.tracks_arrays(want.tracks), G.tracks_arrays(r.tracks)
                for k in wt:
                    assert np.array_equal(gt[k], wt[k]), (got_t, s, k)
                seen["processed"] += int(want.processed)
                if want.skip_reason:
                    seen[want.skip_reason] += 1
                seen["raised"] += int(workers[s].process_every > 1)
        h.poll_status()
        # the run must actually have walked every branch of the state machine
        assert seen["motion"] > S and seen["adaptive"] > 0 and seen["raised"] > 0 and seen["processed"] > 4 * S, seen
    finally:
        h.close()


@pytest.mark.timeout(900)
def test_config5_dense_200_ticks_long_lived_tracks():
    import torch
    from realtime_video_analytics_32streams_b200 import B200IouTracker, TrackerConfig, _native

    S, T = 2, 200
    h = _native.Handle(device=0, max_batch=S, max_anchors=8400, max_candidates=4096, max_dets=1024, max_streams=S,
                       max_tracks=4096)
    try:
        trk = B200IouTracker(TrackerConfig(max_age=30, max_iou_distance=0.5, min_hits=1), handle=h)
        ora = O.IouTracker(30, 0.5, 1)
        scenes = [synth.DenseScene(5 + s, orbit=3.0, jitter=0.7) for s in range(S)]
        names = [f"dense{s}" for s in range(S)]
        lbs = [_native.letterbox_meta(1080, 1920, 640, 640)] * S
        meta = O.letterbox_meta(1080, 1920, 640, 640)
        first_ids = None
        for t in range(T):
            heads = np.stack([sc.head(t) for sc in scenes])
            soa = h.postprocess(torch.from_numpy(heads).to(h.device), lbs, 0.35, 0.5, filter_conf=0.35)
            host = B200IouTracker.soa_to_host(trk.update_batch(names, soa))
            dcount = soa["count"].cpu().numpy()
            for s in range(S):
                dets_o = O.filter_detections(O.postprocess(heads[s][None], meta, 0.35, 0.5), 0.35)
                assert int(dcount[s]) == len(dets_o), (t, s)
                want = G.tracks_arrays(ora.update(names[s], dets_o))
                got = G.tracks_arrays(B200IouTracker.tracks_from_soa(host, s))
                for k, v in got.items():
                    assert np.array_equal(v, want[k]), (t, s, k)
                assert len(want["id"]) > 250
            if t == 0:
                first_ids = set(G.tracks_arrays(B200IouTracker.tracks_from_soa(host, 0))["id"].tolist())
        # "long-lived tracks for the whole 200-frame run": most of the first frame's tracks are still alive, with their hits
        last = G.tracks_arrays(B200IouTracker.tracks_from_soa(host, 0))
        alive = first_ids & set(last["id"].tolist())
        assert len(alive) > 0.8 * len(first_ids), (len(alive), len(first_ids))
        assert int(last["hits"].max()) >= T - 5
        h.poll_status()
    finally:
        h.close()

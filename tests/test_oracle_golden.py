"""The CPU oracle replayed against golden vectors produced by the reference's own functions
(tests/golden/make_golden.py).  CPU-only; this is what pins the oracle."""
import numpy as np
import pytest

from oracle import hotpath as O
from realtime_video_analytics_32streams_b200 import synth

import golden_util as G


def test_preprocess_small_cases_bit_exact():
    z = G.load("preprocess")
    for i in range(int(z["pre_n"][0])):
        seed, h, w, ih, iw, half = z[f"pre{i}_cfg"].tolist()
        frame = z[f"pre{i}_frame"]
        assert np.array_equal(frame, synth.synth_frame(seed, h, w))
        tensor, meta = O.preprocess(frame, (ih, iw), bool(half))
        ref = z[f"pre{i}_tensor"]
        assert tensor.dtype == ref.dtype and tensor.shape == ref.shape
        assert np.array_equal(tensor.view(np.uint8), ref.view(np.uint8)), f"case {i}"
        oh, ow, left, top = z[f"pre{i}_meta"].tolist()
        assert meta["orig_shape"] == (oh, ow) and meta["pad"] == (left, top)
        assert meta["scale"] == float(z[f"pre{i}_scale"][0])


@pytest.mark.parametrize("key", ["1080p_f32", "4k_f32", "720p_f32", "odd_f32", "demo360_f32", "portrait_f32",
                                 "1080p_f16", "4k_f16", "odd_f16"])
def test_preprocess_full_size_digests(key):
    d = G.meta()["preprocess_digests"][key]
    frame = synth.synth_frame(d["seed"], d["h"], d["w"])
    tensor, meta = O.preprocess(frame, (640, 640), d["half"])
    assert G.sha(tensor) == d["sha256"]
    assert meta["scale"] == d["scale"] and list(meta["pad"]) == d["pad"]


def test_preprocess_rknn_u8_variant_bit_exact():
    """a2: the oracle's ``preprocess_u8`` against ``RKNNDetector._preprocess`` itself (detector.py:777-839)."""
    z = G.load("preprocess_rknn")
    for i in range(int(z["rk_n"][0])):
        seed, h, w, ih, iw, nhwc = z[f"rk{i}_cfg"].tolist()
        tensor, meta = O.preprocess_u8(synth.synth_frame(seed, h, w), (ih, iw), bool(nhwc))
        ref = z[f"rk{i}_tensor"]
        assert tensor.dtype == ref.dtype == np.uint8 and tensor.shape == ref.shape
        assert np.array_equal(tensor, ref), f"case {i}"
        oh, ow, left, top = z[f"rk{i}_meta"].tolist()
        assert meta["orig_shape"] == (oh, ow) and meta["pad"] == (left, top)
        assert meta["scale"] == float(z[f"rk{i}_scale"][0])


@pytest.mark.parametrize("key", ["1080p_nhwc", "1080p_nchw", "4k_nhwc", "odd_nchw", "portrait_nhwc"])
def test_preprocess_rknn_full_size_digests(key):
    d = G.meta()["preprocess_rknn_digests"][key]
    tensor, meta = O.preprocess_u8(synth.synth_frame(d["seed"], d["h"], d["w"]), (640, 640), d["nhwc"])
    assert G.sha(tensor) == d["sha256"]
    assert meta["scale"] == d["scale"] and list(meta["pad"]) == d["pad"]


def test_postprocess_cases_bit_exact():
    z = G.load("postprocess")
    for name in z["post_names"].tolist():
        conf, iou, oh, ow, ih, iw = z[f"post_{name}_cfg"].tolist()
        meta = O.letterbox_meta(int(oh), int(ow), int(ih), int(iw))
        classes = z[f"post_{name}_classes"].tolist() or None
        dets = O.postprocess(z[f"post_{name}_head"], meta, conf, iou, classes)
        cls, cf, box = G.dets_arrays(dets)
        assert np.array_equal(cls, z[f"post_{name}_cls"]), name
        assert np.array_equal(cf, z[f"post_{name}_conf"]), name
        assert np.array_equal(box, z[f"post_{name}_box"]), name


def test_postprocess_dense_digest():
    d = G.meta()["postprocess_digests"]["dense_seed5_t0"]
    head = synth.DenseScene(5).head(0)[None]
    dets = O.postprocess(head, O.letterbox_meta(1080, 1920, 640, 640), 0.35, 0.5)
    cls, cf, box = G.dets_arrays(dets)
    assert len(dets) == d["n"]
    assert (G.sha(cls), G.sha(cf), G.sha(box)) == (d["cls"], d["conf"], d["box"])


def test_postprocess_rejects_batched_input_like_reference():
    with pytest.raises(ValueError):
        O.postprocess(np.zeros((2, 84, 100), np.float32), O.letterbox_meta(1080, 1920, 640, 640), 0.3, 0.5)


def test_filter_detections_is_float64():
    # float32(0.45) < 0.45: passes the f32 gate of _postprocess, dropped by filter_detections
    d = [O.Det(0, float(np.float32(0.45)), (0, 0, 1, 1)), O.Det(0, 0.5, (0, 0, 1, 1))]
    assert len(O.filter_detections(d, 0.45)) == 1


def test_tracker_sequences_bit_exact():
    z = G.load("tracker")
    for ci in z["trk_cases"].tolist():
        max_age, thr, min_hits = z[f"trk{ci}_cfg"].tolist()
        trk = O.IouTracker(int(max_age), thr, int(min_hits))
        for fi in range(int(z[f"trk{ci}_n"][0])):
            stream = "cam%d" % int(z[f"trk{ci}_{fi}_stream"][0])
            dets = [O.Det(int(c), float(s), tuple(b.tolist())) for c, s, b in
                    zip(z[f"trk{ci}_{fi}_dcls"], z[f"trk{ci}_{fi}_dconf"], z[f"trk{ci}_{fi}_dbox"])]
            got = G.tracks_arrays(trk.update(stream, dets))
            for k, v in got.items():
                assert np.array_equal(v, z[f"trk{ci}_{fi}_t{k}"]), (ci, fi, k)


def test_motion_filter_sequence():
    z = G.load("filters")
    mf = O.MotionFilter(0.02)
    for t, frame in enumerate(z["motion_frames"]):
        assert mf.should_process(frame) == bool(z["motion_decisions"][t])
        assert np.array_equal(mf.previous_gray, z["motion_grays"][t])


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_apply_roi(name):
    z = G.load("filters")
    polys = G.polys_from(z[f"roi_{name}_polys"], z[f"roi_{name}_sizes"])
    assert np.array_equal(O.apply_roi(z["roi_frame"], polys), z[f"roi_{name}_out"])


def test_downsample():
    z = G.load("filters")
    assert np.array_equal(O.downsample(z["down_in"], 0.5), z["down_out_05"])
    assert np.array_equal(O.downsample(z["down_in"], 0.37), z["down_out_037"])
    src = z["down_in"]
    assert O.downsample(src, 0.9995) is src  # frame_filter.py:54: pass-through


@pytest.mark.parametrize("name", ["1080p", "4k"])
def test_filters_full_size_digests(name):
    d = G.meta()["filters_digests"][name]
    polys = [[tuple(p) for p in poly] for poly in d["polys"]]
    assert polys == synth.synth_polygons(d["seed"], d["h"], d["w"])
    assert G.sha(O.roi_mask((d["h"], d["w"]), polys)) == d["mask_sha256"]
    frame = synth.synth_frame(d["seed"], d["h"], d["w"])
    assert G.sha(O.apply_roi(frame, polys)) == d["roi_sha256"]
    mf = O.MotionFilter()
    mf.should_process(frame)
    assert G.sha(mf.previous_gray) == d["blur_sha256"]


def replay_pipeline(make_worker):
    """Shared driver: replays tests/golden/pipeline.npz through per-stream workers.
    ``make_worker(i, name, spec_kwargs, cfg)`` returns an object with ``process(frame, head)``
    -> (processed, tracks-arrays dict, process_every, idle_frames)."""
    z = G.load("pipeline")
    h, w, ih, iw = z["pipe_cfg"].tolist()
    polys = [[tuple(p) for p in z["pipe_polys"].tolist()]]
    scenes = {0: synth.MotionScene(61, h, w, rect=30, speed=11), 1: synth.MotionScene(62, h, w, static=True)}
    burst = synth.MotionScene(63, h, w, rect=30, speed=11)
    workers = {}
    for i in (0, 1):
        spec = dict(name=f"cam{i}", roi_polygons=polys if i == 0 else None, motion_filter=True, motion_threshold=0.02,
                    downsample_ratio=1.0 if i == 0 else 0.75, adaptive_fps=True, target_fps=25, min_target_fps=5,
                    idle_frame_tolerance=3)
        workers[i] = make_worker(i, spec, dict(conf=0.35, iou=0.5, input_hw=(ih, iw), max_age=3, thr=0.5, min_hits=1))
    for k in range(int(z["pipe_n"][0])):
        i, t, processed, pe, idle = z[f"pipe_{k}_hdr"].tolist()
        awake = i == 0 or 8 <= t <= 12
        n_obj = 5 if (awake and t < 14) else 0
        head = synth.synth_head(7000 + 10 * t + i, 20, 256, n_obj, dup=2, input_hw=(ih, iw))[None]
        frame = (burst if (i == 1 and 8 <= t <= 12) else scenes[i]).frame(t)
        assert [G.sha(frame), G.sha(head)] == z[f"pipe_{k}_sha"].tolist(), "synthetic generator drifted"
        got_processed, tracks, got_pe, got_idle = workers[i].process(frame, head)
        assert (got_processed, got_pe, got_idle) == (bool(processed), pe, idle), (k, i, t)
        for kk, v in tracks.items():
            assert np.array_equal(v, z[f"pipe_{k}_t{kk}"]), (k, i, t, kk)


def test_pipeline_state_machine_matches_reference_worker():
    tracker_box = {}

    class W:
        def __init__(self, i, spec, cfg):
            if "t" not in tracker_box:
                tracker_box["t"] = O.IouTracker(cfg["max_age"], cfg["thr"], cfg["min_hits"])
            self.head = None
            self.w = O.StreamWorker(O.StreamSpec(**spec), lambda tensor, idx: self.head, tracker_box["t"],
                                    cfg["conf"], cfg["iou"], None, cfg["input_hw"])

        def process(self, frame, head):
            self.head = head
            r = self.w.process(frame)
            return r.processed, G.tracks_arrays(r.tracks), self.w.process_every, self.w.idle_frames

    replay_pipeline(W)

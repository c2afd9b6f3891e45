#!/usr/bin/env python
"""Generate the golden vectors in this directory by running the REFERENCE's own code.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports ``realtime_analytics`` from ``/root/reference/src`` (nothing is copied), drives
the reference's ``_TensorRTBaseDetector._preprocess/_postprocess``, ``IouTracker.update``,
``MotionFilter.should_process``, ``apply_roi``, ``downsample`` and
``StreamWorker._process_packet`` on seeded inputs and stores inputs + outputs as small
``.npz`` files.  Full-size cases (1080p / 4K -> 640x640) are stored as SHA-256 digests of the
output bytes; their inputs are regenerated from the seed at test time.

Versions the vectors were produced with are recorded in ``golden_meta.json``.
"""

from __future__ import annotations

import asyncio
import hashlib
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference/src")

import cv2  # noqa: E402
from realtime_analytics.config import DetectorConfig, StreamConfig, TrackerConfig  # noqa: E402
from realtime_analytics.detector import Detection, RKNNDetector, _TensorRTBaseDetector, filter_detections  # noqa: E402
from realtime_analytics.pipeline import StreamHealth, StreamWorker, StreamWorkerContext  # noqa: E402
from realtime_analytics.tracker import IouTracker  # noqa: E402
from realtime_analytics.utils import MotionFilter, MotionFilterConfig, apply_roi, downsample  # noqa: E402
from realtime_analytics.video_stream import FramePacket  # noqa: E402

from realtime_video_analytics_32streams_b200 import synth  # noqa: E402

sys.path.insert(0, os.path.dirname(HERE))
from golden_util import egress_case  # noqa: E402  (seeded inputs shared with the tests)


class StubDetector(_TensorRTBaseDetector):
    """The reference's numpy pre/post path with the model forward replaced by a lookup."""

    def __init__(self, config, input_hw):
        super().__init__(config, input_hw)
        self.head = None
        self.head_fn = None
        self.calls = 0

    def _infer(self, tensor):
        self.calls += 1
        if self.head_fn is not None:
            return self.head_fn(tensor, self.calls)
        return self.head


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def dets_to_arrays(dets):
    return (np.array([d.class_id for d in dets], dtype=np.int64),
            np.array([d.confidence for d in dets], dtype=np.float64),
            np.array([d.bbox_xyxy for d in dets], dtype=np.float64).reshape(-1, 4))


def tracks_to_arrays(tracks):
    return (np.array([t.track_id for t in tracks], dtype=np.int64),
            np.array([t.class_id for t in tracks], dtype=np.int64),
            np.array([t.confidence for t in tracks], dtype=np.float64),
            np.array([t.bbox_xyxy for t in tracks], dtype=np.float64).reshape(-1, 4),
            np.array([t.age for t in tracks], dtype=np.int64),
            np.array([t.hits for t in tracks], dtype=np.int64))


STREAM = StreamConfig(name="s0", url="x")


def gen_preprocess(out):
    cases = [  # (seed, h, w, in_h, in_w, half)
        (11, 54, 96, 32, 32, False), (12, 100, 37, 64, 64, False), (13, 37, 100, 64, 64, False),
        (14, 90, 160, 64, 96, False), (15, 48, 48, 64, 64, False), (16, 135, 240, 64, 64, True),
        (17, 33, 77, 96, 64, False), (18, 20, 20, 64, 64, True), (19, 216, 384, 64, 64, False),
    ]
    for i, (seed, h, w, ih, iw, half) in enumerate(cases):
        det = StubDetector(DetectorConfig(backend="tensorrt", half=half), (ih, iw))
        frame = synth.synth_frame(seed, h, w)
        tensor, meta = det._preprocess(frame)
        out[f"pre{i}_frame"] = frame
        out[f"pre{i}_tensor"] = tensor
        out[f"pre{i}_meta"] = np.array([meta["orig_shape"][0], meta["orig_shape"][1], meta["pad"][0], meta["pad"][1]], dtype=np.int64)
        out[f"pre{i}_scale"] = np.array([meta["scale"]], dtype=np.float64)
        out[f"pre{i}_cfg"] = np.array([seed, h, w, ih, iw, int(half)], dtype=np.int64)
    out["pre_n"] = np.array([len(cases)])
    digests = {}
    for name, seed, h, w in [("1080p", 1000, 1080, 1920), ("4k", 4000, 2160, 3840), ("720p", 720, 720, 1280),
                             ("odd", 77, 1083, 1921), ("demo360", 360, 360, 640), ("portrait", 91, 1920, 1080)]:
        frame = synth.synth_frame(seed, h, w)
        for half in (False, True):
            det = StubDetector(DetectorConfig(backend="tensorrt", half=half), (640, 640))
            tensor, meta = det._preprocess(frame)
            digests[f"{name}_{'f16' if half else 'f32'}"] = {
                "seed": seed, "h": h, "w": w, "half": half, "sha256": sha(tensor),
                "scale": meta["scale"], "pad": list(meta["pad"]),
            }
    return digests


def gen_preprocess_rknn(out):
    """a2: ``RKNNDetector._preprocess`` (detector.py:777-839) -- letterbox that stays BGR uint8, NHWC or NCHW.
    ``RKNNDetector.__init__`` imports the ``rknn`` runtime (absent here), so the instance is built with
    ``object.__new__`` and given exactly the three attributes ``_preprocess`` reads."""

    def make(input_hw, nhwc):
        det = object.__new__(RKNNDetector)
        det.config = DetectorConfig(backend="rknn")
        det.input_hw = input_hw
        det.use_nhwc = nhwc
        return det

    cases = [  # (seed, h, w, in_h, in_w, nhwc)
        (111, 54, 96, 32, 32, True), (112, 100, 37, 64, 64, False), (113, 37, 100, 64, 64, True),
        (114, 90, 160, 64, 96, False), (115, 48, 48, 64, 64, True), (116, 135, 240, 64, 64, False),
        (117, 33, 77, 96, 64, True), (118, 216, 384, 64, 64, True),
    ]
    for i, (seed, h, w, ih, iw, nhwc) in enumerate(cases):
        frame = synth.synth_frame(seed, h, w)
        tensor, meta = make((ih, iw), nhwc)._preprocess(frame)
        out[f"rk{i}_tensor"] = tensor
        out[f"rk{i}_meta"] = np.array([meta["orig_shape"][0], meta["orig_shape"][1], meta["pad"][0], meta["pad"][1]], dtype=np.int64)
        out[f"rk{i}_scale"] = np.array([meta["scale"]], dtype=np.float64)
        out[f"rk{i}_cfg"] = np.array([seed, h, w, ih, iw, int(nhwc)], dtype=np.int64)
    out["rk_n"] = np.array([len(cases)])
    digests = {}
    for name, seed, h, w in [("1080p", 1000, 1080, 1920), ("4k", 4000, 2160, 3840), ("720p", 720, 720, 1280),
                             ("odd", 77, 1083, 1921), ("portrait", 91, 1920, 1080)]:
        frame = synth.synth_frame(seed, h, w)
        for nhwc in (True, False):
            tensor, meta = make((640, 640), nhwc)._preprocess(frame)
            digests[f"{name}_{'nhwc' if nhwc else 'nchw'}"] = {
                "seed": seed, "h": h, "w": w, "nhwc": nhwc, "sha256": sha(tensor), "scale": meta["scale"],
                "pad": list(meta["pad"])}
    return digests


def gen_postprocess(out):
    frame = np.zeros((1080, 1920, 3), np.uint8)
    cases = []

    def run(name, head, conf, iou, classes=None, model_type="yolov8", hw=(1080, 1920), in_hw=(640, 640)):
        det = StubDetector(DetectorConfig(backend="tensorrt", confidence_threshold=conf, iou_threshold=iou,
                                          classes=classes, model_type=model_type), in_hw)
        _, meta = det._preprocess(np.zeros((hw[0], hw[1], 3), np.uint8))
        dets = det._postprocess(head, FramePacket(STREAM, frame, 1, 0.0), meta)
        cls, conf_a, box = dets_to_arrays(dets)
        out[f"post_{name}_head"] = head
        out[f"post_{name}_cls"], out[f"post_{name}_conf"], out[f"post_{name}_box"] = cls, conf_a, box
        out[f"post_{name}_cfg"] = np.array([conf, iou, hw[0], hw[1], in_hw[0], in_hw[1]], dtype=np.float64)
        out[f"post_{name}_classes"] = np.array(classes if classes else [], dtype=np.int64)
        cases.append(name)
        return dets

    run("v8_sparse", synth.synth_head(21, 84, 300, 12)[None], 0.35, 0.5)
    run("v8_dup", synth.synth_head(22, 84, 400, 20, dup=5)[None], 0.35, 0.5)
    run("v8_dup_tight", synth.synth_head(23, 84, 400, 20, dup=5)[None], 0.25, 0.3, hw=(720, 1280))
    run("v8_whitelist", synth.synth_head(24, 84, 300, 30, n_obj_classes=6)[None], 0.35, 0.5, classes=[0, 2, 5])
    run("v5_anchor_major", synth.synth_head(25, 85, 500, 25, dup=3, anchor_major=True)[None], 0.45, 0.45, model_type="yolov5")
    run("v8_empty", synth.synth_head(26, 84, 300, 0)[None], 0.35, 0.5)
    run("c5_single", np.ascontiguousarray(synth.synth_head(27, 84, 200, 15)[:5])[None], 0.5, 0.5)
    run("v8_4k_portrait", synth.synth_head(28, 84, 300, 25, dup=4)[None], 0.35, 0.5, hw=(3840, 2160))
    run("v8_smallnet", synth.synth_head(29, 20, 256, 18, dup=3, input_hw=(320, 416))[None], 0.35, 0.6,
        hw=(600, 800), in_hw=(320, 416))
    # touching / nested boxes around the IoU threshold
    head = synth.synth_head(30, 84, 300, 0)
    k = 0
    for gx in range(6):
        for j in range(4):
            a = 10 + k
            head[:4, a] = (100 + gx * 80, 300, 40 + j * 6, 40 - j * 3)
            head[4, a] = 0.95
            head[5 + (gx % 3), a] = 0.9 - 0.01 * k
            k += 1
    run("v8_nested", head[None], 0.35, 0.5)
    out["post_names"] = np.array(cases)
    # full-size dense case as digests
    dense = synth.DenseScene(5).head(0)[None]
    det = StubDetector(DetectorConfig(backend="tensorrt", confidence_threshold=0.35, iou_threshold=0.5), (640, 640))
    _, meta = det._preprocess(frame)
    dets = det._postprocess(dense, FramePacket(STREAM, frame, 1, 0.0), meta)
    cls, conf_a, box = dets_to_arrays(dets)
    return {"dense_seed5_t0": {"n": len(dets), "cls": sha(cls), "conf": sha(conf_a), "box": sha(box)}}


def gen_tracker(out):
    rng = np.random.default_rng(41)
    names = []
    for ci, (max_age, thr, min_hits) in enumerate([(30, 0.7, 3), (30, 0.5, 1), (2, 0.3, 0)]):
        trk = IouTracker(TrackerConfig(max_age=max_age, max_iou_distance=thr, min_hits=min_hits))
        n_obj = 14
        pos = np.stack([rng.uniform(50, 1800, n_obj), rng.uniform(50, 1000, n_obj),
                        rng.uniform(30, 120, n_obj), rng.uniform(30, 120, n_obj)], 1)
        vel = rng.uniform(-4, 4, (n_obj, 2))
        cls = rng.integers(0, 3, n_obj)
        frames = []
        for t in range(40):
            for si, stream in enumerate(("camA", "camB")):
                dets = []
                if not (t % 11 == 7 and si == 1):  # an empty (skip) frame now and then
                    for o in range(n_obj):
                        if rng.random() < 0.15:
                            continue  # miss
                        cx, cy = pos[o, 0] + vel[o, 0] * t + si * 13, pos[o, 1] + vel[o, 1] * t
                        w, h = pos[o, 2], pos[o, 3]
                        reps = 2 if rng.random() < 0.2 else 1  # duplicate -> multi-match
                        for r in range(reps):
                            j = rng.normal(0, 1.5, 4) if r else np.zeros(4)
                            b = np.array([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2]) + j
                            b = b.astype(np.float32).astype(np.float64)
                            dets.append(Detection(stream, t, int(cls[o]), float(np.float32(rng.uniform(0.4, 1))), tuple(b.tolist())))
                    order = rng.permutation(len(dets))
                    dets = [dets[i] for i in order]
                tracks = trk.update(stream, dets)
                frames.append((si, dets, tracks_to_arrays(tracks)))
        out[f"trk{ci}_cfg"] = np.array([max_age, thr, min_hits], dtype=np.float64)
        out[f"trk{ci}_n"] = np.array([len(frames)])
        for fi, (si, dets, tr) in enumerate(frames):
            c, s, b = dets_to_arrays(dets)
            out[f"trk{ci}_{fi}_stream"] = np.array([si])
            out[f"trk{ci}_{fi}_dcls"], out[f"trk{ci}_{fi}_dconf"], out[f"trk{ci}_{fi}_dbox"] = c, s, b
            for k, arr in zip(("id", "cls", "conf", "box", "age", "hits"), tr):
                out[f"trk{ci}_{fi}_t{k}"] = arr
        names.append(ci)
    out["trk_cases"] = np.array(names)


def gen_filters(out):
    # motion: small moving scene
    sc = synth.MotionScene(51, 72, 128, rect=20, speed=3)
    mf = MotionFilter(MotionFilterConfig(enable=True, threshold=0.02), (72, 128, 3))
    decisions, grays = [], []
    frames = []
    for t in range(8):
        f = sc.frame(t) if t not in (3, 4) else frames[-1].copy()  # two repeated (static) frames
        frames.append(f)
        decisions.append(mf.should_process(f))
        grays.append(mf.previous_gray.copy())
    out["motion_frames"] = np.stack(frames)
    out["motion_decisions"] = np.array(decisions)
    out["motion_grays"] = np.stack(grays)
    # roi: in-bounds and out-of-bounds polygons
    frame = synth.synth_frame(52, 60, 80)
    polys_a = [[(10, 5), (70, 12), (60, 50), (20, 55)], [(0, 0), (15, 0), (0, 15)]]
    polys_b = [[(-20, 10), (50, -15), (100, 40), (30, 80)], [(40, 20), (45, 20), (45, 25)]]
    polys_c = [[(5, 5), (70, 50), (70, 5), (5, 50)]]  # self-intersecting
    for name, polys in (("a", polys_a), ("b", polys_b), ("c", polys_c)):
        out[f"roi_{name}_out"] = apply_roi(frame, polys)
        out[f"roi_{name}_polys"] = np.array([p for poly in polys for p in poly], dtype=np.int64)
        out[f"roi_{name}_sizes"] = np.array([len(p) for p in polys], dtype=np.int64)
    out["roi_frame"] = frame
    out["down_in"] = synth.synth_frame(53, 90, 150)
    out["down_out_05"] = downsample(out["down_in"], 0.5)
    out["down_out_037"] = downsample(out["down_in"], 0.37)
    digests = {}
    for name, seed, h, w in (("1080p", 1000, 1080, 1920), ("4k", 4000, 2160, 3840)):
        polys = synth.synth_polygons(seed, h, w)
        mask = np.zeros((h, w), np.uint8)
        for p in polys:
            cv2.fillPoly(mask, [np.array(p, np.int32)], 255)
        f0 = synth.synth_frame(seed, h, w)
        mf = MotionFilter(MotionFilterConfig(enable=True), f0.shape)
        mf.should_process(f0)
        digests[name] = {"seed": seed, "h": h, "w": w, "polys": polys, "mask_sha256": sha(mask),
                         "roi_sha256": sha(apply_roi(f0, polys)), "blur_sha256": sha(mf.previous_gray)}
    return digests


class _Null:
    def __getattr__(self, name):
        async def _async(*a, **k):
            return None

        def _sync(*a, **k):
            return None

        return _async if name == "send_tracks" else _sync


def gen_pipeline(out):
    """Drive the reference's StreamWorker._process_packet (pipeline.py:143-212) with stub sinks."""
    h, w, in_hw = 108, 192, (64, 64)
    scenes = {"cam0": synth.MotionScene(61, h, w, rect=30, speed=11), "cam1": synth.MotionScene(62, h, w, static=True)}
    burst = synth.MotionScene(63, h, w, rect=30, speed=11)  # cam1 wakes up for a few frames
    det = StubDetector(DetectorConfig(backend="tensorrt", confidence_threshold=0.35, iou_threshold=0.5), in_hw)
    tracker = IouTracker(TrackerConfig(max_age=3, max_iou_distance=0.5, min_hits=1))
    polys = [[(10, 8), (180, 12), (170, 100), (20, 95)]]
    workers, streams = {}, {}
    for i, name in enumerate(scenes):
        sc = StreamConfig(name=name, url="x", roi_polygons=polys if i == 0 else None, motion_filter=True,
                          motion_threshold=0.02, downsample_ratio=1.0 if i == 0 else 0.75, adaptive_fps=True,
                          target_fps=25, min_target_fps=5, idle_frame_tolerance=3)
        ctx = StreamWorkerContext(stream=sc, detector=det, tracker=tracker, kafka=_Null(), metrics=_Null(),
                                  health=StreamHealth(name))
        wk = StreamWorker(ctx)
        wk._maybe_save_snapshot = lambda *a, **k: None
        workers[name], streams[name] = wk, sc
    n_frames = 24
    heads = []
    records = []

    async def drive():
        for t in range(n_frames):
            for i, name in enumerate(scenes):
                awake = name == "cam0" or 8 <= t <= 12
                n_obj = 5 if (awake and t < 14) else 0
                head = synth.synth_head(7000 + 10 * t + i, 20, 256, n_obj, dup=2, input_hw=in_hw)[None]
                det.head = head
                before = det.calls
                frame = (burst if (name == "cam1" and 8 <= t <= 12) else scenes[name]).frame(t)
                await workers[name]._process_packet(FramePacket(streams[name], frame, t, 0.0))
                processed = det.calls != before
                tr = tracks_to_arrays(list(tracker._tracks.get(name, {}).values()))
                records.append((i, t, processed, frame, head, tr, workers[name]._process_every, workers[name]._idle_frames))

    asyncio.run(drive())
    out["pipe_n"] = np.array([len(records)])
    out["pipe_cfg"] = np.array([h, w, in_hw[0], in_hw[1]])
    out["pipe_polys"] = np.array(polys[0], dtype=np.int64)
    for k, (i, t, processed, frame, head, tr, pe, idle) in enumerate(records):
        out[f"pipe_{k}_hdr"] = np.array([i, t, int(processed), pe, idle], dtype=np.int64)
        # frames / heads are regenerated from their seeds at test time; digests guard drift
        out[f"pipe_{k}_sha"] = np.array([sha(frame), sha(head)])
        for kk, arr in zip(("id", "cls", "conf", "box", "age", "hits"), tr):
            out[f"pipe_{k}_t{kk}"] = arr


def gen_egress(out):
    """KafkaSink.send_tracks / _render_frame (sinks/kafka_sink.py:93-149, 200-301) with a stub producer: the payload
    bytes the producer's value_serializer writes, and the image handed to cv2.imencode (captured, the encoder still runs)."""
    from realtime_analytics.config import KafkaSinkConfig
    from realtime_analytics.sinks.kafka_sink import KafkaSink
    from realtime_analytics.tracker import Track

    digests = {}
    for name in ("small", "small_overlap", "hd_plus", "uhd", "uhd_dense", "qhd"):
        frame, ids, cls, conf, box = egress_case(name)
        tracks = [Track(int(i), int(c), float(f), tuple(float(v) for v in b)) for i, c, f, b in zip(ids, cls, conf, box)]
        sink = KafkaSink(KafkaSinkConfig(enabled=True, include_frames=True, frame_quality=75))
        sent, captured = [], []

        class Producer:
            async def send_and_wait(self, topic, payload):
                sent.append(json.dumps(payload).encode("utf-8"))  # the value_serializer of kafka_sink.py:88

        sink._producer = Producer()
        real_imencode = cv2.imencode

        def spy(ext, img, params=None):
            captured.append((ext, img.copy(), list(params or [])))
            return real_imencode(ext, img, params)

        cv2.imencode = spy
        try:
            asyncio.run(sink.send_tracks(f"cam-{name}", 1234, tracks, frame=frame))
        finally:
            cv2.imencode = real_imencode
        assert len(sent) == 1 and len(captured) == 1
        ext, image, params = captured[0]
        body = sent[0]
        cut = body.index(b', "frame_jpeg": ')
        digests[name] = {"image_sha256": sha(image), "image_shape": list(image.shape), "ext": ext, "params": [int(p) for p in params],
                         "body_sha256": hashlib.sha256(body).hexdigest(), "body_len": len(body)}
        out[f"{name}_event"] = np.frombuffer(body[:cut] + b"}", dtype=np.uint8)  # the document without the preview
        if image.size <= 170000:
            out[f"{name}_image"] = image
            out[f"{name}_body"] = np.frombuffer(body, dtype=np.uint8)
    return digests


GENERATORS = (("egress", gen_egress), ("preprocess", gen_preprocess), ("preprocess_rknn", gen_preprocess_rknn), ("postprocess", gen_postprocess),
              ("tracker", gen_tracker), ("filters", gen_filters), ("pipeline", gen_pipeline))


def main():
    """``--only NAME [NAME ...]`` regenerates just those files and merges their digests into golden_meta.json."""
    only = sys.argv[sys.argv.index("--only") + 1:] if "--only" in sys.argv else None
    meta = {"cv2": cv2.__version__, "numpy": np.__version__, "python": sys.version.split()[0],
            "reference": "/root/reference (skygazer42/realtime-video-analytics-32streams)"}
    meta_path = os.path.join(HERE, "golden_meta.json")
    if only and os.path.exists(meta_path):
        with open(meta_path) as fh:
            old = json.load(fh)
        assert (old["cv2"], old["numpy"]) == (meta["cv2"], meta["numpy"]), "library versions changed: regenerate everything"
        meta = old
    for name, fn in GENERATORS:
        if only and name not in only:
            continue
        out = {}
        digests = fn(out)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        if digests:
            meta[f"{name}_digests"] = digests
        print(name, len(out), "arrays", os.path.getsize(os.path.join(HERE, f"{name}.npz")) // 1024, "KiB")
    with open(os.path.join(HERE, "golden_meta.json"), "w") as fh:
        json.dump(meta, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

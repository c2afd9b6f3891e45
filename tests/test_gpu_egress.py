"""Egress on the GPU: ``b200va_resize_area_u8`` and ``b200va_draw_rects`` against the oracle
(oracle/egress.py, pinned against cv2 and the reference's KafkaSink by tests/test_oracle_egress.py), and
``B200KafkaSink`` end to end against the golden vectors captured from the reference's own ``KafkaSink.send_tracks``:
the image handed to the encoder and the published message body, byte for byte."""
import asyncio
import hashlib

import numpy as np
import pytest

import golden_util as G
from oracle import egress as E
from realtime_video_analytics_32streams_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def H():
    import torch
    from realtime_video_analytics_32streams_b200 import _native

    assert torch.cuda.is_available()
    h = _native.Handle(device=0, max_batch=16, max_anchors=256, max_candidates=256, max_dets=64, max_streams=8, max_tracks=64)
    yield h
    h.close()


def cu(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_resize_area_matches_oracle_on_every_path(H):
    rng = np.random.default_rng(3)
    shapes = [(216, 384, 108, 192), (300, 500, 100, 250), (90, 150, 30, 50), (144, 256, 108, 192), (152, 269, 108, 191),
              (100, 100, 37, 53), (217, 333, 108, 165), (64, 64, 64, 32), (35, 50, 7, 10), (110, 200, 105, 192), (80, 120, 80, 120)]
    frames = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w, _, _ in shapes]
    outs = H.resize_area([cu(f) for f in frames], [(nh, nw) for _, _, nh, nw in shapes])  # one call, several geometries
    for f, o, (h, w, nh, nw) in zip(frames, outs, shapes):
        assert np.array_equal(o.cpu().numpy(), E.resize_area(f, nw, nh)), (h, w, nh, nw)
    # a batch that shares one geometry (one launch) and a pitched source
    batch = [rng.integers(0, 256, (120, 200, 3), dtype=np.uint8) for _ in range(5)]
    outs = H.resize_area([cu(f) for f in batch], [(90, 150)] * 5)
    for f, o in zip(batch, outs):
        assert np.array_equal(o.cpu().numpy(), E.resize_area(f, 150, 90))
    wide = cu(rng.integers(0, 256, (60, 130, 3), dtype=np.uint8))
    view = wide[:, 10:110]
    assert np.array_equal(H.resize_area([view], [(30, 50)])[0].cpu().numpy(), E.resize_area(view.cpu().numpy(), 50, 30))
    from realtime_video_analytics_32streams_b200 import _native

    with pytest.raises(_native.B200VAError):  # enlarging is not what INTER_AREA is used for here
        H.resize_area([cu(batch[0])], [(240, 400)])


@pytest.mark.parametrize("hw,new", [((2160, 3840), (1080, 1920)), ((1440, 2560), (1080, 1920)), ((1100, 2000), (1056, 1920))])
def test_resize_area_full_size_matches_cv2_restatement(H, hw, new):
    frame = synth.synth_frame(900 + hw[0], *hw)
    got = H.resize_area([cu(frame)], [new])[0].cpu().numpy()
    assert np.array_equal(got, E.resize_area(frame, new[1], new[0]))


def test_draw_rects_matches_sequential_cv2_semantics(H):
    rng = np.random.default_rng(4)
    for rnd in range(30):
        sizes = [(int(rng.integers(8, 90)), int(rng.integers(8, 120))) for _ in range(int(rng.integers(1, 6)))]
        imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in sizes]
        ops_all = []
        for h, w in sizes:
            ops = []
            for _ in range(int(rng.integers(0, 25))):
                p = [int(rng.integers(-20, w + 20)), int(rng.integers(-20, h + 20)), int(rng.integers(-20, w + 20)), int(rng.integers(-20, h + 20))]
                if rng.random() < 0.15:
                    p[2] = p[0]
                if rng.random() < 0.15:
                    p[3] = p[1]
                ops.append((int(rng.integers(0, 2)), *p, tuple(int(v) for v in rng.integers(0, 256, 3))))
            ops_all.append(ops)
        dev = [cu(i) for i in imgs]
        H.draw_rects(dev, ops_all)
        for img, d, ops in zip(imgs, dev, ops_all):
            want = img.copy()
            E.apply_ops(want, ops)
            assert np.array_equal(d.cpu().numpy(), want), (rnd, img.shape, len(ops))


def test_draw_rects_crowded_tile_and_vector_resize_fallbacks(H):
    """More operations over one tile than the kernel's per-tile list holds (it then walks the whole list), and 2x2
    shrinking of frames the vector kernel cannot take (odd destination width, unaligned rows)."""
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (40, 90, 3), dtype=np.uint8)
    ops = [(int(rng.integers(0, 2)), int(rng.integers(0, 60)), int(rng.integers(0, 25)), int(rng.integers(20, 90)), int(rng.integers(10, 40)),
            tuple(int(v) for v in rng.integers(0, 256, 3))) for _ in range(700)]
    dev = cu(img)
    H.draw_rects([dev], [ops])
    want = img.copy()
    E.apply_ops(want, ops)
    assert np.array_equal(dev.cpu().numpy(), want)
    for h, w in ((64, 130), (64, 136)):  # dst width 65 (odd) / 68 on a source view that starts 3 bytes into a row
        f = rng.integers(0, 256, (h, w + 1, 3), dtype=np.uint8)
        view = cu(f)[:, 1:]
        assert np.array_equal(H.resize_area([view], [(h // 2, w // 2)])[0].cpu().numpy(), E.resize_area(f[:, 1:], w // 2, h // 2))


class _Producer:
    def __init__(self):
        self.sent = []

    async def send_and_wait(self, topic, body):
        self.sent.append((topic, body))


@pytest.mark.parametrize("name", G.EGRESS_CASES)
def test_sink_publishes_the_reference_bytes(H, name, monkeypatch):
    """B200KafkaSink.send_tracks with include_frames: the pre-encode image and the whole message body equal what the
    reference's KafkaSink produced for the same frame and tracks (golden).  The non-overlapping cases must take the GPU
    drawing path; the overlapping ones are replayed in the reference's order."""
    import cv2
    from realtime_video_analytics_32streams_b200 import sinks
    from realtime_video_analytics_32streams_b200.types import Track

    frame, ids, cls, conf, box = G.egress_case(name)
    want = G.meta()["egress_digests"][name]
    cfg = type("Cfg", (), dict(enabled=True, include_frames=True, frame_quality=75, topic="analytics",
                               bootstrap_servers="x", linger_ms=10, max_batch_size=16384))()
    prod = _Producer()
    sink = sinks.B200KafkaSink(cfg, handle=H, producer=prod)
    captured = []
    real = cv2.imencode
    monkeypatch.setattr(cv2, "imencode", lambda ext, img, params=None: (captured.append((ext, img.copy(), list(params or []))), real(ext, img, params))[1])
    tracks = [Track(int(i), int(c), float(f), tuple(float(v) for v in b), 0, 1) for i, c, f, b in zip(ids, cls, conf, box)]
    asyncio.run(sink.send_tracks(f"cam-{name}", 1234, tracks, frame=frame))
    assert len(prod.sent) == 1 and len(captured) >= 1
    ext, image, params = captured[-1]
    assert ext == want["ext"] and [int(p) for p in params] == want["params"]
    assert list(image.shape) == want["image_shape"] and G.sha(image) == want["image_sha256"]
    topic, body = prod.sent[0]
    assert topic == "analytics" and len(body) == want["body_len"] and hashlib.sha256(body).hexdigest() == want["body_sha256"]
    assert (sink._renderer.replayed == 1) == (name in ("small_overlap", "uhd_dense"))
    # the rate limit of kafka_sink.py:151-163: a second event inside the interval goes out without the preview
    sink._frame_send_interval = 3600.0
    asyncio.run(sink.send_tracks(f"cam-{name}", 1235, {"track_id": ids, "cls": cls, "conf": conf, "bbox_xyxy": box}, frame=frame))
    assert b"frame_jpeg" not in prod.sent[1][1] and prod.sent[1][1].startswith(b'{"stream": "cam-')


def test_sink_disabled_or_without_frames(H):
    from realtime_video_analytics_32streams_b200 import sinks

    frame, ids, cls, conf, box = G.egress_case("small")
    cfg = type("Cfg", (), dict(enabled=True, include_frames=False, frame_quality=75, topic="t"))()
    prod = _Producer()
    sink = sinks.B200KafkaSink(cfg, handle=H, producer=prod)
    asyncio.run(sink.send_tracks("cam-small", 1234, {"track_id": ids, "cls": cls, "conf": conf, "bbox_xyxy": box}, frame=frame))
    assert prod.sent[0][1] == bytes(G.load("egress")["small_event"])
    cfg.enabled = False
    asyncio.run(sink.send_tracks("cam-small", 1, [], frame=frame))
    assert len(prod.sent) == 1
    assert [sink._calculate_adaptive_quality(n) for n in (0, 1, 3, 4, 10, 11)] == [E.adaptive_quality(75, n) for n in (0, 1, 3, 4, 10, 11)]

"""Egress on the CPU: the oracle's OpenCV restatements pinned against the installed cv2, the oracle
against the golden vectors made by the reference's own KafkaSink, and ``b200va_tracks_json`` (host-only C code, needs no
GPU) byte for byte against ``json.dumps`` and against the reference's payload."""
import hashlib
import json
import struct

import numpy as np
import pytest

import golden_util as G
from oracle import egress as E
from realtime_video_analytics_32streams_b200 import _native, sinks

cv2 = pytest.importorskip("cv2")


def test_resize_area_restatement_matches_cv2():
    rng = np.random.default_rng(0)
    # whole 2x2 blocks, other whole blocks (3x3, 4x2), fractional ratios incl. the sink's 0.96 / 0.75 shapes scaled down
    for h, w, nh, nw in ((216, 384, 108, 192), (300, 500, 100, 250), (90, 150, 30, 50), (144, 256, 108, 192), (152, 269, 108, 191),
                         (100, 100, 37, 53), (217, 333, 108, 165), (64, 64, 64, 32), (35, 50, 7, 10), (110, 200, 105, 192)):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(E.resize_area(img, nw, nh), cv2.resize(img, (nw, nh), interpolation=cv2.INTER_AREA)), (h, w, nh, nw)


def test_rectangle_restatements_match_cv2():
    rng = np.random.default_rng(1)
    for it in range(1500):
        h, w = int(rng.integers(5, 60)), int(rng.integers(5, 80))
        a = np.zeros((h, w, 3), np.uint8)
        b = a.copy()
        p1 = (int(rng.integers(-15, w + 15)), int(rng.integers(-15, h + 15)))
        p2 = (int(rng.integers(-15, w + 15)), int(rng.integers(-15, h + 15)))
        if it % 7 == 0:
            p2 = (p1[0], p2[1])
        if it % 11 == 0:
            p2 = (p2[0], p1[1])
        col = tuple(int(v) for v in rng.integers(1, 256, 3))
        if it % 2:
            cv2.rectangle(a, p1, p2, col, 2)
            E.draw_rect2(b, p1, p2, col)
        else:
            cv2.rectangle(a, p1, p2, col, -1)
            E.fill_rect(b, p1, p2, col)
        assert np.array_equal(a, b), (it, p1, p2)


def _label_size(label):
    (lw, lh), base = cv2.getTextSize(label, cv2.FONT_HERSHEY_SIMPLEX, 0.5, 2)
    return int(lw), int(lh), int(base)


def _oracle_preview(frame, ids, cls, box):
    h, w = frame.shape[:2]
    sf, nw, nh = E.preview_geometry(h, w)
    img = E.resize_area(frame, nw, nh) if (nh, nw) != (h, w) else frame.copy()
    tl = [{"track_id": int(i), "class_id": int(c), "bbox_xyxy": tuple(b)} for i, c, b in zip(ids, cls, box.tolist())]
    sizes = [_label_size(f"ID {t['track_id']}") for t in tl]
    ops = E.overlay_ops(tl, sf, sizes)
    for k, trk in enumerate(tl):  # strictly in the reference's order: box, background, text per track
        E.apply_ops(img, ops[2 * k:2 * k + 2])
        x1, y1 = ops[2 * k][1], ops[2 * k][2]
        cv2.putText(img, f"ID {trk['track_id']}", (x1, max(0, y1 - 4)), cv2.FONT_HERSHEY_SIMPLEX, 0.5, (255, 255, 255), 2, cv2.LINE_AA)
    return img


@pytest.mark.parametrize("name", ["small", "small_overlap", "hd_plus"])
def test_oracle_preview_matches_the_reference_sink(name):
    """The image the reference's KafkaSink._render_frame hands to cv2.imencode (captured by tests/golden/make_golden.py)."""
    frame, ids, cls, conf, box = G.egress_case(name)
    img = _oracle_preview(frame, ids, cls, box)
    want = G.meta()["egress_digests"][name]
    assert list(img.shape) == want["image_shape"] and G.sha(img) == want["image_sha256"]
    g = G.load("egress")
    if f"{name}_image" in g:
        assert np.array_equal(img, g[f"{name}_image"])


def test_host_side_plan_matches_the_oracle_and_flags_overlaps():
    for name in G.EGRESS_CASES:
        frame, ids, cls, conf, box = G.egress_case(name)
        sf, nw, nh = sinks.preview_geometry(*frame.shape[:2])
        assert (sf, nw, nh) == E.preview_geometry(*frame.shape[:2])
        tl = [{"track_id": int(i), "class_id": int(c), "bbox_xyxy": tuple(b)} for i, c, b in zip(ids, cls, box.tolist())]
        ops, texts, conflict = sinks.overlay_plan(tl, sf, _label_size)
        assert ops == E.overlay_ops(tl, sf, [_label_size(t[0]) for t in texts])
        assert conflict == (name in ("small_overlap", "uhd_dense")), name
    assert [sinks.color_for(c) for c in (0, 1, 79, -1, 2 ** 24, 16777215)] == [E.color_for(c) for c in (0, 1, 79, -1, 2 ** 24, 16777215)]


def test_tracks_json_is_json_dumps_byte_for_byte():
    def check(name, fid, ids, cls, conf, box, url=None):
        tracks = [{"track_id": int(i), "class_id": int(c), "confidence": float(f), "bbox_xyxy": tuple(float(v) for v in b)}
                  for i, c, f, b in zip(ids, cls, conf, box)]
        payload = {"stream": name, "frame_id": fid, "tracks": tracks, "is_temporal": False}
        if url is not None:
            payload["frame_jpeg"] = url
        assert _native.tracks_json(name, fid, ids, cls, conf, box, url) == json.dumps(payload).encode("utf-8")

    check("cam-01", 12, [1, 2], [0, 5], [0.9, 0.5], [[1, 2, 3, 4], [0.1, 0.2, 0.3, 0.4]])
    check("摄像头 \"7\"\\\n\t\x01\x7f😀é/", -3, [], [], [], np.zeros((0, 4)))
    check("s", 2 ** 62, [2 ** 62, -5], [-1, 2 ** 31 - 1], [float("nan"), float("inf")],
          [[float("-inf"), -0.0, 1e16, 1e15], [1e-4, 1e-5, 5e-324, 1.7976931348623157e308]], "data:image/jpeg;base64,AB+/=")
    rng = np.random.default_rng(2)
    vals = [float(np.float32(v)) for v in rng.uniform(-2000, 4000, 8000)]
    vals += [struct.unpack("<d", struct.pack("<Q", int(rng.integers(0, 2 ** 63)) | (int(rng.integers(0, 2)) << 63)))[0] for _ in range(8000)]
    vals += [10.0 ** k for k in range(-30, 31)] + [float(k) for k in range(-50, 50)] + [123456789012345680.0, 9999999999999998.0, 0.30000000000000004]
    vals = np.array(vals[:len(vals) // 4 * 4])
    box = vals.reshape(-1, 4)
    m = box.shape[0]
    check("x", 1, np.arange(m), np.zeros(m, int), vals[:m], box)


@pytest.mark.parametrize("name", G.EGRESS_CASES)
def test_tracks_json_matches_the_reference_payload(name):
    """The bytes the reference's producer serialiser wrote for the same tracks (golden, preview key cut off)."""
    frame, ids, cls, conf, box = G.egress_case(name)
    want = bytes(G.load("egress")[f"{name}_event"])
    assert _native.tracks_json(f"cam-{name}", 1234, ids, cls, conf, box) == want
    assert json.loads(want)["tracks"][0]["track_id"] == int(ids[0])

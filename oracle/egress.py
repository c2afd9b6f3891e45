"""TEST INFRASTRUCTURE -- CPU restatement of the reference's egress formats.

Restates, in NumPy / plain Python, what ``KafkaSink.send_tracks`` and ``KafkaSink._render_frame``
(``/root/reference/src/realtime_analytics/sinks/kafka_sink.py:93-149, 200-310``) do to a frame and its tracks
before the bytes leave the process:

* the event payload and its ``json.dumps`` serialisation (kafka_sink.py:88, 105-134);
* the preview: ``cv2.resize(..., INTER_AREA)`` when the frame is larger than 1920x1080 (:227-232), per track
  ``cv2.rectangle(.., color, 2)`` (:240), the filled label background (:249-255), ``cv2.putText`` (:258-267), then
  ``cv2.imencode`` (:270-296) and base64 (:299-300).

The OpenCV primitives (INTER_AREA, the thickness-2 rectangle, the filled rectangle) are restated from OpenCV 4.x's
``modules/imgproc/src/resize.cpp`` (``ResizeAreaFast_`` / ``ResizeArea_`` / ``computeResizeAreaTab``) and
``drawing.cpp`` (``rectangle`` -> ``PolyLine`` -> ``ThickLine`` -> ``FillConvexPoly`` + ``Circle`` caps); OpenCV is a
third-party dependency of the reference (``opencv-python-headless``, pylock.toml) and is pinned here by
``tests/test_oracle_egress.py`` against the installed ``cv2`` on random inputs, and by the golden vectors
``tests/golden/egress.npz`` generated from the reference's own ``KafkaSink`` (tests/golden/make_golden.py).
Glyph rasterisation (``putText``, Hershey font, LINE_AA) and the JPEG / WebP encoders are NOT restated: the product
calls the same ``cv2`` functions the reference calls for those two steps.

Only tests/, __graft_entry__.smoke() and bench.py's CPU arms may import this module.
"""

from __future__ import annotations

import json
from typing import List, Optional, Sequence, Tuple

import numpy as np


# ---- kafka_sink.py:303-310 -------------------------------------------------------------------
def color_for(class_id: int) -> Tuple[int, int, int]:
    seed = (hash(class_id) & 0xFFFFFF) or 0xFFAA33
    return int(seed & 0xFF), int((seed >> 8) & 0xFF), int((seed >> 16) & 0xFF)


# ---- kafka_sink.py:105-134 + the producer's value_serializer (:88) ----------------------------
def track_payload(stream_name: str, frame_id: int, tracks) -> dict:
    track_list = []
    for t in tracks:
        track_list.append({"track_id": t.track_id, "class_id": t.class_id, "confidence": t.confidence,
                           "bbox_xyxy": t.bbox_xyxy})
    return {"stream": stream_name, "frame_id": frame_id, "tracks": track_list, "is_temporal": False}


def payload_bytes(payload: dict) -> bytes:
    return json.dumps(payload).encode("utf-8")


def adaptive_quality(base_quality: int, n_tracks: int) -> int:  # kafka_sink.py:165-192
    boost = -10 if n_tracks == 0 else (0 if n_tracks <= 3 else (5 if n_tracks <= 10 else 10))
    return max(50, min(95, base_quality + boost))


# ---- cv2.resize(INTER_AREA), 8-bit, shrinking -------------------------------------------------
def _area_tab(ssize: int, dsize: int, scale: float):
    """computeResizeAreaTab (resize.cpp): (si, di, alpha) triples, alpha in float32."""
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1, sx2 = int(np.ceil(fsx1)), int(np.floor(fsx2))
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((sx1 - 1, dx, np.float32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((sx, dx, np.float32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((sx2, dx, np.float32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab


def resize_area(img: np.ndarray, new_w: int, new_h: int) -> np.ndarray:
    """``cv2.resize(img, (new_w, new_h), interpolation=cv2.INTER_AREA)`` for uint8 HxWxC, new <= old in both axes."""
    h, w = img.shape[:2]
    scale_x, scale_y = w / new_w, h / new_h  # double, like inv_scale_x = dsize.width / ssize.width inverted
    iscale_x, iscale_y = int(round(scale_x)), int(round(scale_y))  # saturate_cast<int>
    is_area_fast = abs(scale_x - iscale_x) < 2.220446049250313e-16 and abs(scale_y - iscale_y) < 2.220446049250313e-16
    src = img.astype(np.int64)
    if is_area_fast:
        # ResizeAreaFast_: whole iscale_x x iscale_y blocks; the 2 x 2 case runs the vector kernel (a+b+c+d+2)>>2, every
        # other block size computes saturate_cast<uchar>(sum * (1.f / area)) (float multiply, round half to even)
        blocks = src[:new_h * iscale_y, :new_w * iscale_x].reshape(new_h, iscale_y, new_w, iscale_x, -1).sum(axis=(1, 3))
        if iscale_x == 2 and iscale_y == 2:
            return ((blocks + 2) >> 2).astype(np.uint8)
        scale = np.float32(1.0 / (iscale_x * iscale_y))
        return np.clip(np.rint(blocks.astype(np.float32) * scale), 0, 255).astype(np.uint8)
    # ResizeArea_<uchar, float>: horizontal pass into a float row, vertical accumulation, float arithmetic in table order
    xtab, ytab = _area_tab(w, new_w, scale_x), _area_tab(h, new_h, scale_y)
    cn = img.shape[2] if img.ndim == 3 else 1
    srcf = img.reshape(h, w, cn).astype(np.float32)
    out = np.zeros((new_h, new_w, cn), np.uint8)
    xs = np.array([t[0] for t in xtab])
    xd = np.array([t[1] for t in xtab])
    xa = np.array([t[2] for t in xtab], np.float32)
    # group x-table entries by destination (they are emitted in dx order): position k within its group
    pos = np.zeros(len(xtab), np.int64)
    for i in range(1, len(xtab)):
        pos[i] = pos[i - 1] + 1 if xd[i] == xd[i - 1] else 0
    kmax = int(pos.max()) + 1
    prev_dy, acc = ytab[0][1], np.zeros((new_w, cn), np.float32)
    for sy, dy, beta in ytab:
        buf = np.zeros((new_w, cn), np.float32)
        for k in range(kmax):  # float32 adds in table order, one table position at a time
            sel = pos == k
            buf[xd[sel]] = buf[xd[sel]] + srcf[sy, xs[sel]] * xa[sel, None]
        if dy != prev_dy:
            out[prev_dy] = np.clip(np.rint(acc), 0, 255).astype(np.uint8)
            acc = beta * buf
            prev_dy = dy
        else:
            acc = acc + beta * buf
    out[prev_dy] = np.clip(np.rint(acc), 0, 255).astype(np.uint8)
    return out.reshape((new_h, new_w) + img.shape[2:])


# ---- cv2.rectangle ---------------------------------------------------------------------------
def fill_rect(img: np.ndarray, p1, p2, color) -> None:
    """``cv2.rectangle(img, p1, p2, color, -1)``: the inclusive axis-aligned box, clipped to the image."""
    h, w = img.shape[:2]
    x1, x2 = sorted((int(p1[0]), int(p2[0])))
    y1, y2 = sorted((int(p1[1]), int(p2[1])))
    x1, y1, x2, y2 = max(x1, 0), max(y1, 0), min(x2, w - 1), min(y2, h - 1)
    if x1 <= x2 and y1 <= y2:
        img[y1:y2 + 1, x1:x2 + 1] = color


def draw_rect2(img: np.ndarray, p1, p2, color) -> None:
    """``cv2.rectangle(img, p1, p2, color, 2)``: each of the four edges is a 3-pixel band between its end points
    (ThickLine's convex polygon at half-thickness 1) with round caps of radius 1 (a plus shape) at the corners."""
    x1, x2 = sorted((int(p1[0]), int(p2[0])))
    y1, y2 = sorted((int(p1[1]), int(p2[1])))
    fill_rect(img, (x1, y1 - 1), (x2, y1 + 1), color)
    fill_rect(img, (x1, y2 - 1), (x2, y2 + 1), color)
    fill_rect(img, (x1 - 1, y1), (x1 + 1, y2), color)
    fill_rect(img, (x2 - 1, y1), (x2 + 1, y2), color)


# ---- kafka_sink.py:200-267 without the glyphs -------------------------------------------------
def preview_geometry(h: int, w: int) -> Tuple[float, int, int]:
    """(scale_factor, new_w, new_h) of kafka_sink.py:224-232."""
    if w > 1920 or h > 1080:
        sf = min(1920 / w, 1080 / h)
        return sf, int(w * sf), int(h * sf)
    return 1.0, w, h


def overlay_ops(track_list: Sequence[dict], scale_factor: float, label_sizes: Sequence[Tuple[int, int, int]]):
    """The rectangle operations of kafka_sink.py:235-255 in drawing order:
    (kind, x1, y1, x2, y2, (b, g, r)) with kind 0 = thickness-2 outline, 1 = filled.
    ``label_sizes[i]`` = (label_w, label_h, baseline) of ``cv2.getTextSize(f"ID {track_id}", SIMPLEX, 0.5, 2)``."""
    ops = []
    for trk, (lw, lh, base) in zip(track_list, label_sizes):
        x1, y1, x2, y2 = [int(v * scale_factor) for v in trk["bbox_xyxy"]]
        color = color_for(trk["class_id"])
        ops.append((0, x1, y1, x2, y2, color))
        ops.append((1, x1, max(0, y1 - lh - base - 4), x1 + lw, max(0, y1), color))
    return ops


def apply_ops(img: np.ndarray, ops) -> None:
    for kind, x1, y1, x2, y2, color in ops:
        (fill_rect if kind else draw_rect2)(img, (x1, y1), (x2, y2), color)

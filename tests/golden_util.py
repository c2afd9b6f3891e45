"""Helpers shared by the golden-vector tests."""
import hashlib
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def meta():
    with open(os.path.join(GOLDEN, "golden_meta.json")) as fh:
        return json.load(fh)


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def polys_from(flat, sizes):
    out, k = [], 0
    for n in sizes.tolist():
        out.append([(int(x), int(y)) for x, y in flat[k:k + n].tolist()])
        k += n
    return out


def dets_arrays(dets):
    return (np.array([d.class_id for d in dets], dtype=np.int64),
            np.array([d.confidence for d in dets], dtype=np.float64),
            np.array([d.bbox_xyxy for d in dets], dtype=np.float64).reshape(-1, 4))


def tracks_arrays(tracks):
    return {"id": np.array([t.track_id for t in tracks], dtype=np.int64),
            "cls": np.array([t.class_id for t in tracks], dtype=np.int64),
            "conf": np.array([t.confidence for t in tracks], dtype=np.float64),
            "box": np.array([t.bbox_xyxy for t in tracks], dtype=np.float64).reshape(-1, 4),
            "age": np.array([t.age for t in tracks], dtype=np.int64),
            "hits": np.array([t.hits for t in tracks], dtype=np.int64)}


EGRESS_CASES = ("small", "small_overlap", "hd_plus", "uhd", "uhd_dense", "qhd")


def egress_case(name: str):
    """Seeded inputs of the egress golden cases (shared with tests/test_gpu_egress.py): frame + track rows."""
    cases = {
        # name: (seed, h, w, n_tracks, overlapping)
        "small": (71, 150, 360, 3, False),
        "small_overlap": (72, 120, 200, 9, True),
        "hd_plus": (73, 1100, 2000, 12, False),     # above 1920x1080 -> INTER_AREA with a fractional ratio (0.96)
        "uhd": (74, 2160, 3840, 40, False),         # 4K -> exactly 0.5: whole 2x2 blocks
        "uhd_dense": (75, 2160, 3840, 300, True),
        "qhd": (76, 1440, 2560, 20, False),         # 0.75
    }
    seed, h, w, n, overlap = cases[name]
    rng = np.random.default_rng(seed)
    frame = _synth().synth_frame(seed, h, w)
    ids = rng.permutation(5000)[:n] + 1
    cls = rng.integers(0, 80, n)
    conf = rng.uniform(0.35, 1.0, n).astype(np.float32).astype(np.float64)
    if overlap:
        cx, cy = rng.uniform(0.2 * w, 0.8 * w, n), rng.uniform(0.2 * h, 0.8 * h, n)
        bw, bh = rng.uniform(0.1 * w, 0.4 * w, n), rng.uniform(0.1 * h, 0.4 * h, n)
    else:
        cols = int(np.ceil(np.sqrt(n * w / h)))
        rows = int(np.ceil(n / cols))
        k = np.arange(n)
        cx, cy = (k % cols + 0.5) * w / cols, (k // cols + 0.5) * h / rows + 0.25 * h / rows
        bw, bh = np.full(n, 0.35 * w / cols), np.full(n, 0.3 * h / rows)
        if w / cols < 260:  # labels ("ID 1234" ~ 60 px at preview scale) must not reach the next column's box
            bw = np.full(n, 0.2 * w / cols)
    box = np.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1).astype(np.float32).astype(np.float64)
    box[:, [0, 2]] = np.clip(box[:, [0, 2]], 0, w - 1)
    box[:, [1, 3]] = np.clip(box[:, [1, 3]], 0, h - 1)
    return frame, ids.astype(np.int64), cls.astype(np.int32), conf, box


def _synth():
    from realtime_video_analytics_32streams_b200 import synth

    return synth

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

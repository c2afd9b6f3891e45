"""Dense-scene NMS (k_dense_pairs + k_dense_resolve) against the oracle, with every launch forced onto
those kernels (B200VA_DENSE_IMPL=2): ordinary clusters, the two fall-backs to the single-CTA path (a suppressor list
that overflows, a suppression chain deeper than the Jacobi round limit), chains just inside the limit, frames where
everything is kept (the kept boxes are then ordered by the bitonic sort, not by rank), tiny and empty frames next to
dense ones, class-aware mode, the float64 filter fold, negative thresholds and the Ultralytics-semantics entry."""
import numpy as np
import pytest

from oracle import hotpath as O
from oracle import ultralytics_restate as U
import golden_util as G

pytestmark = pytest.mark.gpu


def cu(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def HD():
    import os

    import torch
    from realtime_video_analytics_32streams_b200 import _native

    assert torch.cuda.is_available()
    os.environ["B200VA_DENSE_IMPL"] = "2"
    try:
        h = _native.Handle(device=0, max_batch=8, max_anchors=8400, max_candidates=4096, max_dets=4096, max_streams=2,
                           max_tracks=64)
    finally:
        del os.environ["B200VA_DENSE_IMPL"]
    yield h
    h.poll_status()
    h.close()


def _head(boxes_xywh, scores, cls=None, n_cls=1, anchors=4096):
    """REF_COMPAT head [5 + n_cls, A]: column 4 = 1.0, one class column holds the confidence."""
    n = len(scores)
    head = np.zeros((5 + n_cls, anchors), np.float32)
    perm = np.random.default_rng(n).permutation(anchors)[:n]  # candidates land in the table in no useful order
    head[:4, perm] = np.asarray(boxes_xywh, np.float32).T
    head[4, perm] = 1.0
    c = np.zeros(n, np.int64) if cls is None else np.asarray(cls)
    head[5 + c, perm] = scores
    return head


def _unique_scores(rng, n, lo=0.4, hi=0.99):
    s = np.unique(rng.uniform(lo, hi, 4 * n + 8).astype(np.float32))
    rng.shuffle(s)
    return s[:n]


def _check(H, heads, fhw, thr, aware=False, filt=None):
    from realtime_video_analytics_32streams_b200 import _native as N

    fh, fw = fhw
    lb = [N.Letterbox(fh, fw, fh, fw, 0, 0, 1.0)] * len(heads)
    meta = {"orig_shape": (fh, fw), "scale": 1.0, "pad": (0, 0)}
    kw = {}
    if filt is not None:
        kw["filter_conf"] = filt
    out = H.postprocess(cu(np.stack(heads)), lb, 0.35, thr, nms_mode=N.NMS_CLASS_AWARE if aware else N.NMS_AGNOSTIC, **kw)
    counts = out["count"].cpu().numpy()
    total = 0
    for b, head in enumerate(heads):
        ref = O.postprocess(head[None], meta, 0.35, thr, class_aware=aware)
        if filt is not None:
            ref = [d for d in ref if d.confidence >= filt]
        m = int(counts[b])
        rc, rf, rb = G.dets_arrays(ref)
        assert m == len(ref), (b, thr, aware, m, len(ref))
        assert np.array_equal(out["cls"][b, :m].cpu().numpy(), rc)
        assert np.array_equal(out["conf"][b, :m].cpu().numpy().astype(np.float64), rf)
        assert np.array_equal(out["bbox_xyxy"][b, :m].cpu().numpy().astype(np.float64), rb, equal_nan=True)
        total += m
    return total


def _clusters(rng, n_obj, dup, fhw, n_cls=1):
    fh, fw = fhw
    boxes, cls = [], []
    for _ in range(n_obj):
        cx, cy = rng.uniform(40, fw - 40), rng.uniform(40, fh - 40)
        w, h = rng.uniform(20, 120, 2)
        c = int(rng.integers(0, n_cls))
        for _ in range(dup):
            boxes.append((cx + rng.normal(0, 1), cy + rng.normal(0, 1), w + rng.normal(0, 1), h + rng.normal(0, 1)))
            cls.append(c if rng.uniform() < 0.8 else int(rng.integers(0, n_cls)))
    return boxes, cls


def test_dense_clusters_mixed_batch(HD):
    rng = np.random.default_rng(1)
    fhw = (1080, 1920)
    heads = []
    for n_obj, dup in ((300, 6), (1, 1), (0, 0), (40, 3), (2, 1), (500, 8), (120, 12), (3, 90)):
        boxes, cls = _clusters(rng, n_obj, dup, fhw, 6)
        n = len(boxes)
        heads.append(_head(boxes, _unique_scores(rng, n), cls, 6) if n else np.zeros((11, 4096), np.float32))
    for thr in (0.5, 0.3):
        for aware in (False, True):
            assert _check(HD, heads, fhw, thr, aware) > 500
    assert _check(HD, heads, fhw, 0.5, filt=0.6) > 150  # filter_detections folded into the emit step
    assert _check(HD, heads, fhw, -0.1) == 7  # a negative threshold: the best box of a frame suppresses every other one


def test_dense_fallback_full_suppressor_lists(HD):
    """One object reported by 200 anchors: the weakest copy has 199 suppressors ahead of it (the list holds 32)."""
    rng = np.random.default_rng(2)
    fhw = (1000, 1000)
    heads = []
    for copies in (200, 34, 33, 32):
        boxes = [(500 + rng.normal(0, 0.5), 500 + rng.normal(0, 0.5), 100, 100) for _ in range(copies)]
        extra, _ = _clusters(rng, 100, 4, fhw)
        boxes += extra
        heads.append(_head(boxes, _unique_scores(rng, len(boxes))))
    assert _check(HD, heads, fhw, 0.5) > 300


def test_dense_suppression_chains(HD):
    """Staircases: box k overlaps box k + 1 above the threshold and box k + 2 below it, scores descending along the
    stairs, so greedy NMS keeps every other box and the dependency depth is the length of the staircase: 40 and 60
    (inside the Jacobi round limit), 70 and 600 (beyond it: the frame is handed to the sequential path)."""
    rng = np.random.default_rng(3)
    fhw = (1080, 4000)
    heads = []
    for length in (40, 60, 63, 64, 65, 70, 600):
        boxes = [(30 + 6.0 * k, 500, 24, 40) for k in range(length)]  # IoU 18/30 = 0.6, then 12/36 = 0.33
        scores = np.sort(_unique_scores(rng, length))[::-1].copy()
        extra, _ = _clusters(rng, 50, 4, (400, 4000))
        boxes += extra
        scores = np.concatenate([scores, _unique_scores(rng, len(extra), 0.36, 0.399)])
        heads.append(_head(boxes, scores))
    kept = _check(HD, heads, fhw, 0.5)
    assert kept >= sum((n + 1) // 2 for n in (40, 60, 63, 64, 65, 70, 600))


def test_dense_nan_boxes(HD):
    """NaN box coordinates (a NaN in the head's box rows): `iou <= thr` is False for them, so they suppress and are
    suppressed like the reference's NumPy expression says; the overlap pre-test must keep such pairs."""
    rng = np.random.default_rng(6)
    fhw = (1080, 1920)
    heads = []
    for k in range(3):
        boxes, _ = _clusters(rng, 150, 5, fhw)
        boxes = np.asarray(boxes, np.float32)
        scores = _unique_scores(rng, len(boxes))
        bad = rng.permutation(len(boxes))[:6]
        boxes[bad[:3], k] = np.nan          # one coordinate
        boxes[bad[3:], :] = np.nan          # the whole box
        scores[bad[0]] = np.float32(0.995)  # one of them leads the frame, the others sit in the middle
        heads.append(_head(boxes, scores))
    with np.errstate(invalid="ignore"):
        _check(HD, heads, fhw, 0.5)


def test_nan_boxes_single_cta_paths():
    """The same NaN rule on the single-CTA kernels: the small-frame path, the general path and its kept-box grid
    (B200VA_DENSE_IMPL=1 keeps dense frames on k_sort_nms; the grid is dropped for frames with a NaN coordinate)."""
    import os

    from realtime_video_analytics_32streams_b200 import _native

    os.environ["B200VA_DENSE_IMPL"] = "1"
    try:
        h = _native.Handle(device=0, max_batch=4, max_anchors=8400, max_candidates=4096, max_dets=4096, max_streams=2,
                           max_tracks=64)
    finally:
        del os.environ["B200VA_DENSE_IMPL"]
    try:
        rng = np.random.default_rng(7)
        fhw = (1080, 1920)
        for n_obj, dup, lead in ((20, 4, True), (150, 5, False), (150, 5, True), (150, 5, False), (30, 3, False)):
            boxes, _ = _clusters(rng, n_obj, dup, fhw)
            boxes = np.asarray(boxes, np.float32)
            scores = _unique_scores(rng, len(boxes))
            bad = rng.permutation(len(boxes))[:4]
            boxes[bad[:2], 2] = np.nan
            boxes[bad[2:], :] = np.nan
            if lead:
                scores[bad[0]] = np.float32(0.995)
            else:
                scores[bad] = np.minimum(scores[bad], np.float32(0.5))
            with np.errstate(invalid="ignore"):
                _check(h, [_head(boxes, scores)], fhw, 0.5)
        h.poll_status()
    finally:
        h.close()


def test_dense_everything_kept(HD):
    """Disjoint boxes: 3000 candidates, 3000 detections (ordered by the bitonic sort), then 500 (ordered by rank)."""
    rng = np.random.default_rng(4)
    fhw = (2160, 3840)
    heads = []
    for n in (3000, 500, 513, 512):
        k = np.arange(n)
        boxes = np.stack([20 + (k % 75) * 50.0, 20 + (k // 75) * 50.0, np.full(n, 30.0), np.full(n, 30.0)], 1)
        heads.append(_head(boxes, _unique_scores(rng, n)))
    assert _check(HD, heads, fhw, 0.5) == 3000 + 500 + 513 + 512


def test_dense_score_ties(HD):
    """Equal scores: the higher candidate (anchor) index comes first; checked against the single-CTA path
    of an ordinary handle, which the oracle-pinned tie test covers."""
    from realtime_video_analytics_32streams_b200 import _native as N

    rng = np.random.default_rng(5)
    fhw = (1080, 1920)
    boxes, _ = _clusters(rng, 200, 6, fhw)
    scores = (np.round(rng.uniform(0.4, 0.99, len(boxes)) * 16) / 16).astype(np.float32)
    head = _head(boxes, scores)
    lb = [N.Letterbox(1080, 1920, 1080, 1920, 0, 0, 1.0)]
    plain = N.Handle(device=0, max_batch=1, max_anchors=8400, max_candidates=4096, max_dets=4096, max_streams=1, max_tracks=64)
    try:
        a = plain.postprocess(cu(head[None]), lb, 0.35, 0.5)
        b = HD.postprocess(cu(head[None]), lb, 0.35, 0.5)
        m = int(a["count"].cpu()[0])
        assert m == int(b["count"].cpu()[0]) and m > 150
        for k in ("cls", "conf", "bbox_xyxy"):
            assert np.array_equal(a[k][0, :m].cpu().numpy(), b[k][0, :m].cpu().numpy()), k
    finally:
        plain.close()


def test_dense_ultralytics_semantics(HD):
    from test_gpu_ultralytics import _compare, v8_head

    in_hw, A = (640, 640), 8400
    heads = [v8_head(500 + s, 10, A, in_hw, 300, 6) for s in range(2)] + [v8_head(77, 10, A, in_hw, 20, 3)]
    fhw = [(1080, 1920)] * 3
    assert _compare(HD, heads, fhw, in_hw, conf_thr=0.3, iou_thr=0.45, max_det=100) == 220
    assert _compare(HD, heads, fhw, in_hw, conf_thr=0.3, iou_thr=0.45, max_det=1000) > 500
    assert _compare(HD, heads, fhw, in_hw, conf_thr=0.3, iou_thr=0.45, agnostic=True, max_det=1000) > 500
    ties = [v8_head(300 + s, 80, A, in_hw, 60, 6, ties=True) for s in range(4)]
    assert _compare(HD, ties, [(1080, 1920), (1920, 1080), (723, 1001), (640, 640)], in_hw, conf_thr=0.25, iou_thr=0.45) > 100

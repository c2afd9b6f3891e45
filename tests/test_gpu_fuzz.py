"""Randomised GPU-vs-oracle comparisons (seeded, deterministic): shapes and inputs chosen to hit the
rarely taken paths -- tracker conflicts / chunk boundaries / candidate overflow, NMS clusters around
the IoU threshold, letterbox extremes, motion borders."""
import numpy as np
import pytest

import golden_util as G
from oracle import hotpath as O
from realtime_video_analytics_32streams_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def H():
    import torch
    from realtime_video_analytics_32streams_b200 import _native

    assert torch.cuda.is_available()
    h = _native.Handle(device=0, max_batch=16, max_anchors=4096, max_candidates=2048, max_dets=512, max_streams=16,
                       max_tracks=1024)
    yield h
    h.close()


def cu(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _tracker_run(H, rng, n_frames, gen, max_age, thr, min_hits, f64):
    from realtime_video_analytics_32streams_b200 import B200IouTracker, TrackerConfig

    H.tracker_set_next_id(1)
    for s in range(4):
        H.tracker_reset(s)
    trk = B200IouTracker(TrackerConfig(max_age=max_age, max_iou_distance=thr, min_hits=min_hits), handle=H)
    ora = O.IouTracker(max_age, thr, min_hits)
    md = 512
    for t in range(n_frames):
        names, rows_all = ["a", "b", "c"], []
        for s in range(3):
            rows_all.append(gen(rng, t, s))
        if f64:
            box = np.zeros((3, md, 4), np.float64)
            conf = np.zeros((3, md), np.float64)
        else:
            box = np.zeros((3, md, 4), np.float32)
            conf = np.zeros((3, md), np.float32)
        cl = np.zeros((3, md), np.int32)
        cnt = np.zeros((3,), np.int32)
        want = []
        for s, rows in enumerate(rows_all):
            cnt[s] = len(rows)
            for i, (c, f, b) in enumerate(rows):
                box[s, i], conf[s, i], cl[s, i] = b, f, c
            dets = [O.Det(int(cl[s, i]), float(conf[s, i]), tuple(float(v) for v in box[s, i])) for i in range(len(rows))]
            want.append(G.tracks_arrays(ora.update(names[s], dets)))
        soa = {"bbox_xyxy": cu(box), "conf": cu(conf), "cls": cu(cl), "count": cu(cnt)}
        out = H.tracker_update([trk.slot_of(n) for n in names], soa, max_age, min_hits, thr, f64=f64)
        host = B200IouTracker.soa_to_host(out)
        for s in range(3):
            got = G.tracks_arrays(B200IouTracker.tracks_from_soa(host, s))
            for k, v in got.items():
                assert np.array_equal(v, want[s][k]), (t, s, k, len(rows_all[s]))
    H.poll_status()


def test_tracker_conflict_storm(H):
    """Heavy overlap: clusters of same-class boxes, duplicates of duplicates, low thresholds -- most
    detections are 'conflicted' (shared candidate tracks and detection-detection edges)."""
    for seed, thr, n_cl, per, f64 in ((1, 0.3, 6, 5, False), (2, 0.1, 4, 9, True), (3, 0.5, 10, 3, False), (4, 0.05, 3, 14, False)):
        rng = np.random.default_rng(seed)
        centers = rng.uniform(100, 900, (3, n_cl, 2))

        def gen(rng, t, s):
            rows = []
            for c in range(n_cl):
                cx, cy = centers[s, c] + rng.normal(0, 3, 2) + t * 2
                for _ in range(int(rng.integers(0, per + 1))):
                    j = rng.normal(0, 6, 4)
                    b = np.array([cx - 40, cy - 30, cx + 40, cy + 30]) + j
                    rows.append((int(rng.integers(0, 2)), np.float32(rng.uniform(0.3, 1)), b.astype(np.float32)))
            order = rng.permutation(len(rows))
            return [rows[i] for i in order]

        _tracker_run(H, rng, 12, gen, max_age=2, thr=thr, min_hits=0, f64=f64)


def test_tracker_many_detections_cross_chunks_and_overflow(H):
    """> 64 detections per frame (several shared-memory chunks), one big pile that exceeds the
    per-detection candidate slots (exact fallback path), and exact duplicates (IoU ties)."""
    rng = np.random.default_rng(7)
    grid = np.stack(np.meshgrid(np.arange(14), np.arange(10)), -1).reshape(-1, 2).astype(np.float64)

    def gen(rng, t, s):
        rows = []
        for k, (gx, gy) in enumerate(grid):
            if rng.random() < 0.1:
                continue
            cx, cy = 60 + gx * 120 + rng.normal(0, 2), 50 + gy * 100 + rng.normal(0, 2)
            b = np.array([cx - 35, cy - 28, cx + 35, cy + 28], np.float32)
            rows.append((k % 3, np.float32(rng.uniform(0.3, 1)), b))
            if rng.random() < 0.15:
                rows.append((k % 3, np.float32(rng.uniform(0.3, 1)), b.copy()))  # exact duplicate: IoU 1.0 tie
        if s == 1:  # a pile of 12 near-identical boxes: more candidates than slots
            for _ in range(12):
                b = np.array([400, 400, 480, 470], np.float32) + rng.normal(0, 1.5, 4).astype(np.float32)
                rows.append((1, np.float32(rng.uniform(0.3, 1)), b))
        order = rng.permutation(len(rows))
        return [rows[i] for i in order]

    _tracker_run(H, rng, 8, gen, max_age=3, thr=0.4, min_hits=1, f64=False)


def test_tracker_working_table_in_global_scratch(monkeypatch):
    """The tracker's working table lives in shared memory only when `live tracks + detections` fits the rows the launch
    was given; otherwise the same code runs on the stream's global scratch.  Force that path (8 shared rows) through the
    conflict-heavy and the multi-chunk scenarios, and let a default handle (512 rows, then adapted from the need the
    kernel reports) cross the limit in both directions: results never change."""
    from realtime_video_analytics_32streams_b200 import _native

    def storm(rng, t, s, centers=np.random.default_rng(5).uniform(100, 900, (3, 5, 2))):
        rows = []
        for c in range(5):
            cx, cy = centers[s, c] + rng.normal(0, 3, 2) + t * 2
            for _ in range(int(rng.integers(0, 7))):
                b = np.array([cx - 40, cy - 30, cx + 40, cy + 30]) + rng.normal(0, 6, 4)
                rows.append((int(rng.integers(0, 2)), np.float32(rng.uniform(0.3, 1)), b.astype(np.float32)))
        return [rows[i] for i in rng.permutation(len(rows))]

    def crowd(rng, t, s):
        # the population swells to ~450 boxes (need ~900 > 512 rows) and shrinks again
        n = (40, 450, 450, 450, 60, 30, 30, 420, 20, 20)[t % 10]
        k = np.arange(n)
        cx, cy = 30 + (k % 30) * 60 + rng.normal(0, 1.5, n), 30 + (k // 30) * 60 + rng.normal(0, 1.5, n)
        return [(int(i % 3), np.float32(rng.uniform(0.3, 1)), np.array([x - 20, y - 20, x + 20, y + 20], np.float32))
                for i, x, y in zip(k, cx, cy)]

    monkeypatch.setenv("B200VA_TRK_SMEM_TRACKS", "8")
    h = _native.Handle(device=0, max_batch=4, max_anchors=256, max_candidates=1024, max_dets=512, max_streams=4, max_tracks=1024)
    try:
        _tracker_run(h, np.random.default_rng(11), 10, storm, max_age=2, thr=0.3, min_hits=0, f64=False)
        _tracker_run(h, np.random.default_rng(12), 10, crowd, max_age=1, thr=0.4, min_hits=1, f64=True)
    finally:
        h.close()
    monkeypatch.delenv("B200VA_TRK_SMEM_TRACKS")
    h = _native.Handle(device=0, max_batch=4, max_anchors=256, max_candidates=1024, max_dets=512, max_streams=4, max_tracks=2048)
    try:
        _tracker_run(h, np.random.default_rng(13), 30, crowd, max_age=1, thr=0.4, min_hits=1, f64=False)
    finally:
        h.close()


def test_tracker_capacity_flag(H):
    from realtime_video_analytics_32streams_b200 import _native as N

    H.poll_status()
    H.tracker_reset(9)
    n = 400
    for t in range(4):  # 4 x 400 non-overlapping new boxes > max_tracks = 1024
        box = np.zeros((1, 512, 4), np.float32)
        xs = (np.arange(n) % 40) * 30.0 + t * 2000
        ys = (np.arange(n) // 40) * 30.0
        box[0, :n] = np.stack([xs, ys, xs + 10, ys + 10], 1)
        soa = {"bbox_xyxy": cu(box), "conf": cu(np.full((1, 512), 0.9, np.float32)), "cls": cu(np.zeros((1, 512), np.int32)),
               "count": cu(np.array([n], np.int32))}
        out = H.tracker_update([9], soa, 30, 0, 0.5)
    assert int(out["count"].cpu()[0]) == 1024
    with pytest.raises(N.B200VAError) as e:
        H.poll_status()
    assert e.value.status == N.ERR_CAPACITY
    H.tracker_reset(9)


def test_nms_clusters_near_threshold(H):
    """Boxes built so that many pairwise IoUs sit within a few ulps of the threshold."""
    from realtime_video_analytics_32streams_b200 import _native as N

    rng = np.random.default_rng(11)
    lb = N.Letterbox(1000, 1000, 1000, 1000, 0, 0, 1.0)
    meta = {"orig_shape": (1000, 1000), "scale": 1.0, "pad": (0, 0)}
    for it in range(6):
        n = int(rng.integers(50, 900))
        base = rng.uniform(50, 900, (n // 8 + 1, 2))
        head = np.zeros((5, 1024), np.float32)
        k = 0
        for bx, by in base:
            w, h = rng.uniform(30, 80, 2)
            for j in range(8):
                if k >= n:
                    break
                # shifting a w x h box by dx gives IoU (w-dx)/(w+dx): dx = w/3 -> exactly 0.5
                dx = w / 3 * (j % 4) + rng.choice([0, 0, 1e-4, -1e-4])
                head[:4, k] = (bx + dx, by + (j // 4) * h * 0.2, w, h)
                k += 1
        scores = np.unique(rng.uniform(0.4, 0.99, 4 * n).astype(np.float32))
        rng.shuffle(scores)
        head[4, :k] = scores[:k]
        for thr in (0.5, 0.45):
            ref = O.postprocess(head[None], meta, 0.35, thr)
            out = H.postprocess(cu(head[None]), [lb], 0.35, thr)
            m = int(out["count"].cpu()[0])
            rc, rf, rb = G.dets_arrays(ref)
            assert m == len(ref), (it, thr)
            assert np.array_equal(out["conf"][0, :m].cpu().numpy().astype(np.float64), rf)
            assert np.array_equal(out["bbox_xyxy"][0, :m].cpu().numpy().astype(np.float64), rb)


@pytest.mark.parametrize("dense_impl", ["1", "0"])
def test_nms_grid_paths(H, dense_impl, monkeypatch):
    """Frames with more than 256 candidates: mixes of tiny boxes crowding one cell (the grid's cell lists overflow; the
    suppressor lists of the pair kernel overflow), frame-sized boxes (too many cells: overflow list), ordinary clusters,
    zero-area boxes and boxes on the frame border, class-agnostic and class-aware, several thresholds and frame shapes.
    dense_impl "1": the single-CTA kernel with the kept-box grid (k_sort_nms<true>, also the fall-back of the dense
    kernels); "0": k_dense_pairs + k_dense_resolve from the second launch on (the default)."""
    from realtime_video_analytics_32streams_b200 import _native as N

    monkeypatch.setenv("B200VA_DENSE_IMPL", dense_impl)
    H = N.Handle(device=0, max_batch=2, max_anchors=4096, max_candidates=2048, max_dets=2048, max_streams=2, max_tracks=64)
    monkeypatch.delenv("B200VA_DENSE_IMPL")
    rng = np.random.default_rng(29)
    for it, (fh, fw) in enumerate([(1080, 1920), (1000, 1000), (2160, 3840), (360, 640), (1920, 1080), (90, 4000)]):
        lb = N.Letterbox(fh, fw, fh, fw, 0, 0, 1.0)
        meta = {"orig_shape": (fh, fw), "scale": 1.0, "pad": (0, 0)}
        n = int(rng.integers(300, 1900))
        C = 4 + 1 + 6
        head = np.zeros((C, 2048), np.float32)
        kinds = rng.choice(5, n, p=[0.45, 0.25, 0.1, 0.1, 0.1])
        cx = rng.uniform(0, fw, n)
        cy = rng.uniform(0, fh, n)
        w = rng.uniform(8, 0.08 * fw, n)
        h = rng.uniform(8, 0.12 * fh, n)
        crowd = kinds == 1  # a lattice of tiny boxes inside one grid cell
        cx[crowd] = 0.3 * fw + rng.integers(0, 12, crowd.sum()) * 7.0
        cy[crowd] = 0.3 * fh + rng.integers(0, 10, crowd.sum()) * 6.0
        w[crowd], h[crowd] = rng.uniform(3, 9, crowd.sum()), rng.uniform(3, 8, crowd.sum())
        huge = kinds == 2
        w[huge], h[huge] = rng.uniform(0.5 * fw, 1.5 * fw, huge.sum()), rng.uniform(0.5 * fh, 1.5 * fh, huge.sum())
        flat = kinds == 3
        w[flat] = 0.0  # zero area: IoU 0 with everything
        edge = kinds == 4
        cx[edge] = rng.choice([0.0, fw - 1.0, fw + 50.0], edge.sum())
        head[0, :n], head[1, :n], head[2, :n], head[3, :n] = cx, cy, w, h
        scores = np.unique(rng.uniform(0.36, 0.999, 4 * n).astype(np.float32))
        rng.shuffle(scores)
        head[4, :n] = 1.0  # objectness; class scores below are the confidences
        cls = rng.integers(0, 6, n)
        head[5 + cls, np.arange(n)] = scores[:n]
        for thr in (0.5, 0.2, 0.0):
            for aware in (False, True):
                ref = O.postprocess(head[None], meta, 0.35, thr, class_aware=aware)
                out = H.postprocess(cu(head[None]), [lb], 0.35, thr, nms_mode=N.NMS_CLASS_AWARE if aware else N.NMS_AGNOSTIC)
                m = int(out["count"].cpu()[0])
                rc, rf, rb = G.dets_arrays(ref)
                assert m == len(ref), (it, thr, aware, m, len(ref))
                assert np.array_equal(out["cls"][0, :m].cpu().numpy(), rc)
                assert np.array_equal(out["conf"][0, :m].cpu().numpy().astype(np.float64), rf)
                assert np.array_equal(out["bbox_xyxy"][0, :m].cpu().numpy().astype(np.float64), rb)
    H.poll_status()
    H.close()


def test_letterbox_random_geometries(H):
    from realtime_video_analytics_32streams_b200 import _native as N

    rng = np.random.default_rng(13)
    for it in range(40):
        h, w = int(rng.integers(2, 700)), int(rng.integers(2, 900))
        ih, iw = int(rng.integers(8, 260)), int(rng.integers(8, 300))
        m = O.letterbox_meta(h, w, ih, iw)
        if m["new_wh"][0] < 1 or m["new_wh"][1] < 1:
            continue
        frame = synth.synth_frame(2000 + it, h, w)
        fmt, half = ((N.OUT_F32_RGB_NCHW, False), (N.OUT_F16_RGB_NCHW, True))[it % 2]
        out, metas = H.preprocess([cu(frame)], (ih, iw), fmt)
        ref, meta = O.preprocess(frame, (ih, iw), half)
        assert np.array_equal(out.cpu().numpy().view(np.uint8), ref.view(np.uint8)), (h, w, ih, iw)
        assert metas[0].as_meta() == meta


def test_motion_random_sizes_and_masks(H):
    import torch

    rng = np.random.default_rng(17)
    for it in range(12):
        h, w = int(rng.integers(3, 400)), int(rng.integers(3, 1100))
        if it % 3 == 0:
            w = 16 * int(rng.integers(1, 60))  # the aligned fast path, incl. widths that are not multiples of 256
        f0, f1 = synth.synth_frame(3000 + it, h, w), synth.synth_frame(4000 + it, h, w)
        f1[:, : w // 2] = f0[:, : w // 2]
        polys = [[(1, 1), (w - 2, 1), (w - 2, h - 2), (w // 2, h // 2), (1, h - 2)]] if (it % 2 and h > 6 and w > 6) else None
        mask = H.roi_rasterize(polys, h, w) if polys else None
        a = torch.empty((h, w), dtype=torch.uint8, device="cuda")
        b = torch.empty((h, w), dtype=torch.uint8, device="cuda")
        H.motion([cu(f0)], [None], [a], [mask] if polys else None)
        c = H.motion([cu(f1)], [a], [b], [mask] if polys else None).cpu().numpy()
        mf = O.MotionFilter()
        mf.should_process(O.apply_roi(f0, polys) if polys else f0)
        assert np.array_equal(a.cpu().numpy(), mf.previous_gray), (h, w)
        mf.should_process(O.apply_roi(f1, polys) if polys else f1)
        assert np.array_equal(b.cpu().numpy(), mf.previous_gray), (h, w)
        assert int(c[0]) == mf.last_count


def test_second_handle_with_smaller_capacities_does_not_break_the_first():
    """Kernel attributes (dynamic shared-memory limit) are per function, not per handle: creating a small handle must
    not lower what a big handle launches with."""
    import torch
    from realtime_video_analytics_32streams_b200 import _native

    big = _native.Handle(device=0, max_batch=2, max_anchors=8400, max_candidates=4096, max_dets=1024, max_streams=2, max_tracks=4096)
    small = None
    try:
        heads = np.stack([synth.DenseScene(77 + s).head(0) for s in range(2)])
        lbs = [_native.letterbox_meta(1080, 1920, 640, 640)] * 2
        want = [len(O.filter_detections(O.postprocess(heads[s][None], O.letterbox_meta(1080, 1920, 640, 640), 0.35, 0.5), 0.35))
                for s in range(2)]
        for rnd in range(3):
            if rnd == 1:
                small = _native.Handle(device=0, max_batch=1, max_anchors=256, max_candidates=64, max_dets=16, max_streams=1, max_tracks=32)
            dets = big.postprocess(torch.from_numpy(heads).cuda(), lbs, 0.35, 0.5, filter_conf=0.35)
            trk = big.tracker_update([0, 1], dets, 30, 1, 0.5)
            assert dets["count"].cpu().tolist() == want and trk["count"].cpu().tolist() == want
            big.poll_status()
    finally:
        big.close()
        if small is not None:
            small.close()

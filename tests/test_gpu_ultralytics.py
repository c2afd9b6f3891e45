"""Ultralytics-semantics entry points (b200va_letterbox_meta_ultralytics / b200va_preprocess_geom /
b200va_postprocess_ultralytics) against oracle/ultralytics_restate.py, bit-exact.  The oracle's NMS step and
float32 box arithmetic are pinned against torchvision / torch CPU (tests/test_oracle_ultralytics.py); against
ultralytics itself parity is unpinned (not installed)."""
import numpy as np
import pytest

from oracle import ultralytics_restate as U
from realtime_video_analytics_32streams_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def H():
    import torch
    from realtime_video_analytics_32streams_b200 import _native

    assert torch.cuda.is_available()
    h = _native.Handle(device=0, max_batch=16, max_anchors=8400, max_candidates=4096, max_dets=1024, max_streams=4,
                       max_tracks=64)
    yield h
    h.poll_status()
    h.close()


def cu(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def v8_head(seed, nc, anchors, in_hw, n_obj, dup, ties=False):
    """Decoded YOLOv8 head [4 + nc, A]: xywh in network-input pixels, class scores in rows 4.."""
    rng = np.random.default_rng(seed)
    h = np.empty((4 + nc, anchors), dtype=np.float32)
    h[0] = rng.uniform(0, in_hw[1], anchors)
    h[1] = rng.uniform(0, in_hw[0], anchors)
    h[2] = rng.uniform(4, 80, anchors)
    h[3] = rng.uniform(4, 80, anchors)
    h[4:] = rng.uniform(0, 0.05, (nc, anchors))
    pick = rng.permutation(anchors)[:n_obj * dup]
    for o in range(n_obj):
        cx, cy = rng.uniform(40, in_hw[1] - 40), rng.uniform(40, in_hw[0] - 40)
        w, hh = rng.uniform(20, 120), rng.uniform(20, 120)
        c = int(rng.integers(0, nc))
        for d in range(dup):
            a = pick[o * dup + d]
            h[:4, a] = (cx + rng.normal(0, 1.5), cy + rng.normal(0, 1.5), w + rng.normal(0, 1.5), hh + rng.normal(0, 1.5))
            s = rng.uniform(0.3, 1.0)
            h[4 + c, a] = np.round(s * 8) / 8 if ties else s
    return h


@pytest.mark.parametrize("auto", [False, True])
def test_preprocess_geom_matches_letterbox_restatement(H, auto):
    from realtime_video_analytics_32streams_b200 import _native as N

    shapes = [(1080, 1920), (2160, 3840), (720, 1280), (1083, 1921), (360, 640), (1920, 1080), (723, 1001), (480, 854)]
    for half, fmt in ((False, N.OUT_F32_RGB_NCHW), (True, N.OUT_F16_RGB_NCHW)):
        for i, (h, w) in enumerate(shapes):
            frame = synth.synth_frame(800 + i, h, w)
            ref, g = U.preprocess(frame, (640, 640), auto=auto, stride=32, half=half)
            m, oh, ow = N.letterbox_meta_ultralytics(h, w, 640, 640, auto, 32)
            assert (oh, ow) == g["out_hw"] == tuple(ref.shape[2:])
            out = H.preprocess_geom([cu(frame)], [m], (oh, ow), fmt)
            got = out.cpu().numpy()
            assert got.shape == ref.shape and got.dtype == ref.dtype
            assert np.array_equal(got.view(np.uint8), ref.view(np.uint8)), (h, w, auto, half)


def _compare(H, heads, frame_hw, in_hw, **kw):
    out = H.postprocess_ultralytics(cu(np.stack(heads)), frame_hw, in_hw, **kw)
    counts = out["count"].cpu().numpy()
    okw = dict(conf_thres=kw.get("conf_thr", 0.25), iou_thres=kw.get("iou_thr", 0.45), classes=kw.get("classes"),
               agnostic=kw.get("agnostic", False), max_det=kw.get("max_det", 300))
    total = 0
    for b, head in enumerate(heads):
        want = U.postprocess(head, in_hw, frame_hw[b], **okw)
        n = int(counts[b])
        assert n == len(want), (b, n, len(want))
        assert out["cls"][b, :n].cpu().tolist() == [c for c, _, _ in want]
        assert np.array_equal(out["conf"][b, :n].cpu().numpy(), np.array([s for _, s, _ in want], dtype=np.float32))
        assert np.array_equal(out["bbox_xyxy"][b, :n].cpu().numpy().reshape(-1, 4),
                              np.array([bx for _, _, bx in want], dtype=np.float32).reshape(-1, 4))
        total += n
    return total


def test_postprocess_ultralytics_rect_input(H):
    in_hw, A = (384, 640), 5040  # 1080p through LetterBox(auto=True): 48x80 + 24x40 + 12x20 anchors
    heads = [v8_head(100 + s, 80, A, in_hw, 40, 5) for s in range(6)]
    frame_hw = [(1080, 1920)] * 4 + [(2160, 3840), (360, 640)]
    assert _compare(H, heads, frame_hw, in_hw, conf_thr=0.25, iou_thr=0.45) > 100
    assert _compare(H, heads, frame_hw, in_hw, conf_thr=0.5, iou_thr=0.7) > 50


def test_postprocess_ultralytics_square_ties_and_modes(H):
    in_hw, A = (640, 640), 8400
    heads = [v8_head(300 + s, 80, A, in_hw, 60, 6, ties=True) for s in range(4)]
    frame_hw = [(1080, 1920), (1920, 1080), (723, 1001), (640, 640)]
    assert _compare(H, heads, frame_hw, in_hw, conf_thr=0.25, iou_thr=0.45) > 100
    assert _compare(H, heads, frame_hw, in_hw, conf_thr=0.25, iou_thr=0.45, agnostic=True) > 100
    assert _compare(H, heads, frame_hw, in_hw, conf_thr=0.25, iou_thr=0.45, classes=[0, 3, 17, 42, 79]) > 0
    assert _compare(H, heads, frame_hw, in_hw, conf_thr=0.25, iou_thr=0.45, max_det=7) == 28
    assert _compare(H, heads, frame_hw, in_hw, conf_thr=0.375, iou_thr=0.5) > 50  # 0.375 is a tie value: strict `>`


def test_postprocess_ultralytics_dense_and_anchor_major(H):
    in_hw, A = (640, 640), 8400
    heads = [v8_head(500 + s, 10, A, in_hw, 300, 6) for s in range(2)]
    assert _compare(H, heads, [(1080, 1920)] * 2, in_hw, conf_thr=0.3, iou_thr=0.45, max_det=100) == 200
    assert _compare(H, heads, [(1080, 1920)] * 2, in_hw, conf_thr=0.3, iou_thr=0.45, max_det=1000) > 500
    # anchor-major layout [B, A, C]
    out_cm = H.postprocess_ultralytics(cu(np.stack(heads)), [(1080, 1920)] * 2, in_hw, 0.3, 0.45)
    out_am = H.postprocess_ultralytics(cu(np.stack([h.T for h in heads])), [(1080, 1920)] * 2, in_hw, 0.3, 0.45)
    for k in ("count", "cls", "conf", "bbox_xyxy"):
        assert np.array_equal(out_cm[k].cpu().numpy(), out_am[k].cpu().numpy()), k


def test_postprocess_ultralytics_edge_cases(H):
    in_hw, A = (640, 640), 8400
    empty = np.zeros((84, A), dtype=np.float32)
    nan = v8_head(9, 80, A, in_hw, 10, 2)
    nan[10, ::7] = np.nan  # a NaN score: torch.max lands on it and `> conf` drops the anchor
    degenerate = v8_head(10, 80, A, in_hw, 10, 3)
    degenerate[2:4, :] = 0.0  # zero-area boxes: 0 / 0 IoU is NaN and never suppresses
    _compare(H, [empty, nan, degenerate], [(1080, 1920)] * 3, in_hw, conf_thr=0.25, iou_thr=0.45)


def test_ultralytics_detector_predict_and_batch(H):
    """B200UltralyticsDetector.predict / predict_batch = LetterBox(auto) -> infer -> NMS -> scale_boxes."""
    import torch
    from realtime_video_analytics_32streams_b200 import (B200UltralyticsDetector, DetectorConfig, FramePacket,
                                                         StreamConfig)

    frames = [synth.synth_frame(60 + i, 1080, 1920) for i in range(3)]
    heads = [v8_head(700 + i, 80, 5040, (384, 640), 30, 4) for i in range(3)]
    seen = []

    def infer(tensor):
        seen.append(tuple(tensor.shape))
        return torch.from_numpy(np.stack(heads[:tensor.shape[0]])).cuda()

    cfg = DetectorConfig(confidence_threshold=0.3, iou_threshold=0.5, classes=None)
    det = B200UltralyticsDetector(cfg, input_hw=(640, 640), infer=infer, handle=H)
    pkt = [FramePacket(StreamConfig(name=f"s{i}"), frames[i], 7 + i, 0.0) for i in range(3)]
    one = det.predict(pkt[0])
    assert seen[-1] == (1, 3, 384, 640)  # rect padding: 1080p -> 384 x 640
    want = U.postprocess(heads[0], (384, 640), (1080, 1920), 0.3, 0.5)
    assert [(d.class_id, d.confidence, d.bbox_xyxy) for d in one] == want
    assert one[0].stream_name == "s0" and one[0].frame_id == 7
    ref, _ = U.preprocess(frames[0], (640, 640), auto=True)
    tensor, meta = det._preprocess(frames[0])
    assert np.array_equal(tensor.cpu().numpy().view(np.uint8), ref.view(np.uint8)) and meta["in_shape"] == (384, 640)
    many = det.predict_batch(pkt)
    assert seen[-1] == (3, 3, 384, 640)
    for i in range(3):
        want = U.postprocess(heads[i], (384, 640), (1080, 1920), 0.3, 0.5)
        assert [(d.class_id, d.confidence, d.bbox_xyxy) for d in many[i]] == want

"""Pins oracle/ultralytics_restate.py where installed software can pin it (torchvision.ops.nms, torch CPU
float32 arithmetic: tests/golden/ultralytics.npz) and states the LetterBox geometry it restates.  The
composition is "parity unpinned": ultralytics itself is not available (see the oracle's docstring)."""
import numpy as np
import pytest

import golden_util as G
from oracle import ultralytics_restate as U


def test_nms_tv_matches_torchvision_golden():
    z = G.load("ultralytics")
    for k in range(int(z["nms_n"][0])):
        b, s = z[f"nms{k}_boxes"], z[f"nms{k}_scores"]
        for t, thr in enumerate((0.45, 0.5, 0.7)):
            keep = U.nms_tv(b, s, thr)
            assert keep.tolist() == z[f"nms{k}_{t}_keep"].tolist(), (k, thr)


def test_nms_tv_matches_installed_torchvision():
    torchvision = pytest.importorskip("torchvision")
    import torch

    rng = np.random.default_rng(7)
    for n in (0, 3, 200, 900):
        c = rng.uniform(0, 300, (n, 2))
        wh = rng.uniform(2, 120, (n, 2))
        b = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
        s = (np.round(rng.uniform(0, 1, n) * 32) / 32).astype(np.float32)
        for thr in (0.3, 0.45, 0.6):
            want = torchvision.ops.nms(torch.from_numpy(b), torch.from_numpy(s), thr).numpy()
            assert U.nms_tv(b, s, thr).tolist() == want.tolist()


def test_class_shift_and_scale_boxes_match_torch_cpu():
    z = G.load("ultralytics")
    b, cls = z["shift_boxes"], z["shift_cls"]
    got = (b + (cls.astype(np.float32) * np.float32(U.MAX_WH))[:, None]).astype(np.float32)
    assert np.array_equal(got.view(np.uint32), z["shift_out"].view(np.uint32))
    for k in range(int(z["scale_n"][0])):
        h1, w1, h0, w0 = z[f"scale{k}_shapes"].tolist()
        got = U.scale_boxes((h1, w1), b, (h0, w0))
        assert np.array_equal(got.view(np.uint32), z[f"scale{k}_out"].view(np.uint32)), k


@pytest.mark.parametrize("hw,auto,want", [
    ((1080, 1920), False, ((640, 360), (0, 140), (0, 140), (640, 640))),
    ((1080, 1920), True, ((640, 360), (0, 12), (0, 12), (384, 640))),
    ((2160, 3840), True, ((640, 360), (0, 12), (0, 12), (384, 640))),
    ((1920, 1080), True, ((360, 640), (12, 0), (12, 0), (640, 384))),
    ((720, 1280), False, ((640, 360), (0, 140), (0, 140), (640, 640))),
    ((723, 1001), False, ((640, 462), (0, 89), (0, 89), (640, 640))),   # round(): the reference truncates to 461? no: 462.24 -> 462
    ((1083, 1921), True, ((640, 361), (0, 11), (0, 12), (384, 640))),   # odd padding: round(11.5 -+ 0.1) = 11 / 12
    ((360, 640), True, ((640, 360), (0, 12), (0, 12), (384, 640))),
])
def test_letterbox_geometry(hw, auto, want):
    g = U.letterbox_geometry(*hw, (640, 640), auto=auto, stride=32)
    assert (g["new_wh"], g["pad"], g["pad_rb"], g["out_hw"]) == want


def test_native_geometry_matches_oracle():
    """The host-only C helper (no GPU needed) against the restatement, over many shapes."""
    import ctypes as C
    from realtime_video_analytics_32streams_b200 import _native as N

    rng = np.random.default_rng(3)
    shapes = [(1080, 1920), (2160, 3840), (720, 1280), (1, 1), (5, 3000), (3000, 5), (641, 641), (639, 640)]
    shapes += [tuple(int(v) for v in rng.integers(8, 4000, 2)) for _ in range(300)]
    for h, w in shapes:
        for auto in (False, True):
            g = U.letterbox_geometry(h, w, (640, 640), auto=auto, stride=32)
            if min(g["new_wh"]) <= 0:
                continue
            m, oh, ow = N.letterbox_meta_ultralytics(h, w, 640, 640, auto, 32)
            assert (m.new_w, m.new_h) == g["new_wh"] and (m.pad_left, m.pad_top) == g["pad"], (h, w, auto)
            assert (oh, ow) == g["out_hw"] and m.scale == g["ratio"]


def test_postprocess_statement_small():
    """A hand-checkable case: two overlapping boxes of one class, one of another class on top of them."""
    pred = np.zeros((4 + 3, 6), dtype=np.float32)
    pred[:4, 0] = (100, 100, 40, 40)
    pred[:4, 1] = (102, 101, 40, 40)   # same class as 0, IoU high -> suppressed
    pred[:4, 2] = (100, 100, 40, 40)   # other class: survives (class shift)
    pred[:4, 3] = (300, 200, 20, 60)
    pred[4, 0], pred[4, 1], pred[5, 2], pred[6, 3] = 0.9, 0.8, 0.85, 0.25
    got = U.postprocess(pred, (384, 640), (1080, 1920), conf_thres=0.25, iou_thres=0.45)
    assert [(c, round(s, 2)) for c, s, _ in got] == [(0, 0.9), (1, 0.85)]  # 0.25 is not > 0.25
    x1, y1, x2, y2 = got[0][2]
    assert (x1, y1, x2, y2) == (240.0, 204.0, 360.0, 324.0)  # (80 - 0) / (1/3), (80 - 12) / (1/3) ... in float32
    agn = U.postprocess(pred, (384, 640), (1080, 1920), conf_thres=0.25, iou_thres=0.45, agnostic=True)
    assert [c for c, _, _ in agn] == [0]
    capped = U.postprocess(pred, (384, 640), (1080, 1920), conf_thres=0.1, iou_thres=0.45, max_det=2)
    assert len(capped) == 2

"""CPU restatement of what ``YOLO(...).predict(frame, conf, iou, classes)`` does around the model
forward -- the call ``UltralyticsDetector.predict`` makes (detector.py:147-155 of the reference).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

PARITY UNPINNED against ultralytics itself: ``ultralytics==8.3.209`` (pylock.toml:1432-1433) is a
third-party dependency that is neither vendored under /root/reference nor installed here, and
there is no network.  This file restates its published algorithm:

* ``LetterBox.__call__``          (ultralytics/data/augment.py)   -> ``letterbox_geometry`` / ``preprocess``
* ``ops.non_max_suppression``     (ultralytics/utils/ops.py)      -> ``non_max_suppression``
* ``ops.scale_boxes`` / ``clip_boxes``                            -> ``scale_boxes``
* ``torchvision.ops.nms`` (csrc/ops/cpu/nms_kernel.cpp)           -> ``nms_tv``

What IS pinned: ``nms_tv`` against the installed ``torchvision.ops.nms`` and the float32 box
arithmetic of ``scale_boxes`` / the class shift against torch CPU tensors
(tests/golden/make_golden_ultralytics.py -> tests/golden/ultralytics.npz), and the resize against
cv2 (the same INTER_LINEAR statement as the reference path).  Device-dependent details follow torch
CPU where torch CPU and CUDA differ (true division by the scalar, float32 IoU compared with the
double threshold) except the input normalisation, where the GPU deployment computes
``im * float32(1/255)`` -- bit-identical to the reference's NumPy path.
"""

from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import cv_restate as cvr

F32 = np.float32
MAX_WH = 7680  # ops.non_max_suppression: class offset in pixels


def letterbox_geometry(h: int, w: int, new_shape=(640, 640), auto: bool = False, stride: int = 32) -> dict:
    """``LetterBox.__call__`` geometry (scaleup=True, center=True, scale_fill=False)."""
    r = min(new_shape[0] / h, new_shape[1] / w)
    new_unpad = int(round(w * r)), int(round(h * r))  # (width, height); Python round = half to even
    dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
    if auto:
        dw, dh = int(np.mod(dw, stride)), int(np.mod(dh, stride))
    dw /= 2
    dh /= 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return {"orig_shape": (h, w), "ratio": r, "new_wh": new_unpad, "pad": (left, top), "pad_rb": (right, bottom),
            "out_hw": (new_unpad[1] + top + bottom, new_unpad[0] + left + right)}


def preprocess(frame: np.ndarray, new_shape=(640, 640), auto: bool = False, stride: int = 32, half: bool = False,
               backend: str = "numpy") -> Tuple[np.ndarray, dict]:
    """``BasePredictor.preprocess``: LetterBox -> BGR2RGB -> CHW -> float -> / 255 (as torch CUDA computes it:
    a multiply by float32(1/255); see the module docstring)."""
    h, w = frame.shape[:2]
    g = letterbox_geometry(h, w, new_shape, auto, stride)
    new_w, new_h = g["new_wh"]
    if (w, h) != (new_w, new_h):
        if backend == "cv2":
            import cv2

            frame = cv2.resize(frame, (new_w, new_h), interpolation=cv2.INTER_LINEAR)
        else:
            frame = cvr.resize_linear_u8(frame, new_w, new_h)
    out_h, out_w = g["out_hw"]
    left, top = g["pad"]
    canvas = np.full((out_h, out_w, 3), 114, dtype=np.uint8)
    canvas[top:top + new_h, left:left + new_w] = frame
    rgb = canvas[:, :, ::-1]
    dtype = np.float16 if half else np.float32
    image = rgb.astype(dtype) * dtype(1.0 / 255.0)
    return np.ascontiguousarray(np.transpose(image, (2, 0, 1)))[None], g


def nms_tv(boxes: np.ndarray, scores: np.ndarray, iou_threshold: float) -> np.ndarray:
    """``torchvision.ops.nms`` CPU kernel: stable descending sort, float32 areas / intersections, the float32
    IoU compared with the double threshold, ``>`` suppresses; NaN never does."""
    boxes = np.asarray(boxes, dtype=F32)
    n = len(boxes)
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    order = np.argsort(-np.asarray(scores, dtype=F32), kind="stable")
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    suppressed = np.zeros(n, dtype=bool)
    keep: List[int] = []
    thr = float(iou_threshold)
    with np.errstate(invalid="ignore", divide="ignore"):
        for _i in range(n):
            i = order[_i]
            if suppressed[i]:
                continue
            keep.append(int(i))
            rest = order[_i + 1:]
            xx1 = np.maximum(x1[i], x1[rest])
            yy1 = np.maximum(y1[i], y1[rest])
            xx2 = np.minimum(x2[i], x2[rest])
            yy2 = np.minimum(y2[i], y2[rest])
            w = np.maximum(F32(0), xx2 - xx1)
            h = np.maximum(F32(0), yy2 - yy1)
            inter = w * h
            ovr = inter / (areas[i] + areas[rest] - inter)  # float32
            suppressed[rest[ovr.astype(np.float64) > thr]] = True
    return np.asarray(keep, dtype=np.int64)


def xywh2xyxy(b: np.ndarray) -> np.ndarray:
    """``ops.xywh2xyxy``: ``wh = x[..., 2:] / 2; xy - wh; xy + wh`` (float32)."""
    b = b.astype(F32, copy=False)
    half = b[:, 2:4] / F32(2)
    return np.concatenate([b[:, :2] - half, b[:, :2] + half], axis=1).astype(F32)


def non_max_suppression(pred: np.ndarray, conf_thres: float = 0.25, iou_thres: float = 0.45,
                        classes: Optional[Sequence[int]] = None, agnostic: bool = False, max_det: int = 300):
    """``ops.non_max_suppression`` for ONE image, ``pred`` = ``[4 + nc, A]`` (multi_label=False, no masks).
    Returns (boxes xyxy in network-input pixels [K,4] f32, conf [K] f32, cls [K] int64, anchor index [K])."""
    pred = np.asarray(pred, dtype=F32)
    x = pred.T  # [A, 4 + nc]
    scores = x[:, 4:]
    with np.errstate(invalid="ignore"):
        xc = scores.max(axis=1) > F32(conf_thres)
    idx = np.nonzero(xc)[0]
    x = x[idx]
    if len(x) == 0:
        z = np.zeros((0,), dtype=np.int64)
        return np.zeros((0, 4), dtype=F32), np.zeros((0,), dtype=F32), z, z
    box = xywh2xyxy(x[:, :4])
    cls_scores = x[:, 4:]
    j = np.argmax(cls_scores, axis=1)  # torch.max(dim): first maximal index
    conf = cls_scores[np.arange(len(x)), j]
    keep_mask = conf > F32(conf_thres)
    if classes is not None and len(classes):
        keep_mask &= np.isin(j, np.asarray(list(classes)))
    box, conf, j, idx = box[keep_mask], conf[keep_mask], j[keep_mask], idx[keep_mask]
    if len(box) == 0:
        z = np.zeros((0,), dtype=np.int64)
        return np.zeros((0, 4), dtype=F32), np.zeros((0,), dtype=F32), z, z
    c = j.astype(F32) * F32(0 if agnostic else MAX_WH)
    shifted = (box + c[:, None]).astype(F32)
    keep = nms_tv(shifted, conf, iou_thres)[:max_det]
    return box[keep], conf[keep], j[keep].astype(np.int64), idx[keep]


def scale_boxes(img1_shape, boxes: np.ndarray, img0_shape) -> np.ndarray:
    """``ops.scale_boxes`` + ``clip_boxes`` (ratio_pad=None, padding=True), float32, torch CPU semantics."""
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad = (round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1), round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1))
    b = boxes.astype(F32).copy()
    b[:, 0] -= F32(pad[0])
    b[:, 1] -= F32(pad[1])
    b[:, 2] -= F32(pad[0])
    b[:, 3] -= F32(pad[1])
    b /= F32(gain)
    b[:, 0] = np.clip(b[:, 0], F32(0), F32(img0_shape[1]))
    b[:, 1] = np.clip(b[:, 1], F32(0), F32(img0_shape[0]))
    b[:, 2] = np.clip(b[:, 2], F32(0), F32(img0_shape[1]))
    b[:, 3] = np.clip(b[:, 3], F32(0), F32(img0_shape[0]))
    return b


def postprocess(pred: np.ndarray, in_hw, orig_hw, conf_thres: float = 0.25, iou_thres: float = 0.45,
                classes: Optional[Sequence[int]] = None, agnostic: bool = False, max_det: int = 300):
    """``DetectionPredictor.postprocess`` for one image: NMS, then boxes back to the original frame.
    Returns a list of (class_id, confidence, (x1, y1, x2, y2)) in keep order -- the ``boxes`` rows that
    ``UltralyticsDetector.predict`` (detector.py:163-177) turns into ``Detection`` objects."""
    box, conf, cls, _ = non_max_suppression(pred, conf_thres, iou_thres, classes, agnostic, max_det)
    box = scale_boxes(in_hw, box, orig_hw)
    return [(int(c), float(s), tuple(float(v) for v in b)) for c, s, b in zip(cls, conf, box)]

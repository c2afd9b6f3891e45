"""Seeded synthetic workloads for the five BASELINE.json configurations.

NumPy only.  Shared by ``bench.py``, the parity tests and the golden-vector script so that
every arm sees byte-identical inputs.  Nothing here touches the GPU.
"""

from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np


def synth_frame(seed: int, h: int = 1080, w: int = 1920) -> np.ndarray:
    """Uniform-random BGR frame, ``[h, w, 3]`` uint8 (configs 1-3, 5)."""
    return np.random.default_rng(seed).integers(0, 256, size=(h, w, 3), dtype=np.uint8)


class MotionScene:
    """Config 4 frame source: static seeded background, three moving 200x200 rectangles
    (8 px/frame) and per-pixel noise in [-3, 3]; ``static=True`` drops the rectangles so the
    motion ratio stays under the gate and the skip / adaptive-FPS path is exercised."""

    def __init__(self, seed: int, h: int = 2160, w: int = 3840, static: bool = False,
                 rect: int = 200, speed: int = 8, n_rect: int = 3):
        rng = np.random.default_rng(seed)
        self.h, self.w, self.static = h, w, static
        # smooth-ish background so the Gaussian does not average everything to grey
        coarse = rng.integers(0, 256, size=(max(h // 40, 1) + 1, max(w // 40, 1) + 1, 3), dtype=np.uint8)
        self.background = np.repeat(np.repeat(coarse, 40, axis=0), 40, axis=1)[:h, :w].copy()
        self.rect = min(rect, max(h // 4, 1), max(w // 4, 1))
        self.speed = speed
        self.pos = np.stack([rng.integers(0, max(w - self.rect, 1), n_rect),
                             rng.integers(0, max(h - self.rect, 1), n_rect)], axis=1).astype(np.int64)
        self.vel = np.array([[1, 0], [0, 1], [1, 1]], dtype=np.int64)[:n_rect] * speed
        self.colors = rng.integers(0, 256, size=(n_rect, 3), dtype=np.uint8)
        self.seed = seed

    def frame(self, t: int) -> np.ndarray:
        img = self.background.astype(np.int16)
        if not self.static:
            for k in range(len(self.pos)):
                x = int((self.pos[k, 0] + t * self.vel[k, 0]) % max(self.w - self.rect, 1))
                y = int((self.pos[k, 1] + t * self.vel[k, 1]) % max(self.h - self.rect, 1))
                img[y:y + self.rect, x:x + self.rect] = self.colors[k].astype(np.int16)
        noise_rng = np.random.default_rng(self.seed * 7919 + 17 * t + 1)  # stateless: frame(t) is a pure function
        img += noise_rng.integers(-3, 4, size=img.shape, dtype=np.int16)
        return np.clip(img, 0, 255).astype(np.uint8)


def synth_polygons(seed: int, h: int, w: int) -> List[List[Tuple[int, int]]]:
    """One convex hexagon + one triangle, all vertices in-bounds (config 4 ROI)."""
    rng = np.random.default_rng(seed)
    cx, cy = w * (0.35 + 0.3 * rng.random()), h * (0.35 + 0.3 * rng.random())
    rx, ry = w * (0.2 + 0.1 * rng.random()), h * (0.2 + 0.1 * rng.random())
    ang = np.sort(rng.random(6) * 0.6 + np.arange(6)) * (2 * np.pi / 6)
    hexagon = [(int(np.clip(cx + rx * np.cos(a), 0, w - 1)), int(np.clip(cy + ry * np.sin(a), 0, h - 1))) for a in ang]
    tri = [(int(rng.integers(0, w)), int(rng.integers(0, h))) for _ in range(3)]
    return [hexagon, tri]


def _unique_scores(rng, n: int, lo: float, hi: float) -> np.ndarray:
    """n distinct float32 values in [lo, hi): oracle tie order is unspecified, avoid ties."""
    vals = np.unique(rng.uniform(lo, hi, size=2 * n + 8).astype(np.float32))
    rng.shuffle(vals)
    assert len(vals) >= n
    return vals[:n]


def synth_head(seed: int, n_classes_plus4: int = 84, n_anchors: int = 8400, n_objects: int = 10,
               dup: int = 1, n_obj_classes: Optional[int] = None, input_hw=(640, 640),
               anchor_major: bool = False, jitter: float = 1.0, centers: Optional[np.ndarray] = None,
               obj_cls: Optional[np.ndarray] = None, return_centers: bool = False):
    """Decoded-head tensor ``[C, A]`` (channel-major, YOLOv8 export layout) or ``[A, C]``
    (anchor-major, YOLOv5 layout) with ``n_objects x dup`` above-threshold candidates.

    REF_COMPAT scoring (detector.py:294-305) multiplies column 4 into columns 5.., so the
    generator puts an "objectness" in column 4 and one strong class column per object; the
    final products are made pairwise distinct.  Background rows stay below 0.05 * 0.05.
    """
    rng = np.random.default_rng(seed)
    c, a = n_classes_plus4, n_anchors
    in_h, in_w = input_hw
    ncls = c - 5
    head = np.empty((a, c), dtype=np.float32)
    head[:, 0] = rng.uniform(0, in_w, a)
    head[:, 1] = rng.uniform(0, in_h, a)
    head[:, 2] = rng.uniform(4, in_w / 4, a)
    head[:, 3] = rng.uniform(4, in_h / 4, a)
    head[:, 4:] = rng.uniform(0, 0.05, size=(a, c - 4))
    n_cand = n_objects * dup
    if n_cand:
        if centers is None:
            centers = np.stack([rng.uniform(0.05 * in_w, 0.95 * in_w, n_objects),
                                rng.uniform(0.25 * in_h, 0.75 * in_h, n_objects),
                                rng.uniform(12, 60, n_objects), rng.uniform(12, 60, n_objects)], axis=1)
        anchors = rng.permutation(a)[:n_cand]
        drawn_cls = rng.integers(0, n_obj_classes or ncls, n_objects)
        obj_cls = drawn_cls if obj_cls is None else np.asarray(obj_cls)
        boxes = np.repeat(centers, dup, axis=0) + np.concatenate(
            [np.zeros((n_objects, 1, 4)), rng.normal(0, jitter, size=(n_objects, max(dup - 1, 0), 4))], axis=1
        ).reshape(n_cand, 4)
        head[anchors, :4] = boxes.astype(np.float32)
        objness = rng.uniform(0.9, 1.0, n_cand).astype(np.float32)
        head[anchors, 4] = objness
        # choose the class probability so the products are all different float32 values
        target = _unique_scores(rng, n_cand, 0.46, 0.99)
        cls_p = (target / objness).astype(np.float32)
        cols = 5 + np.repeat(obj_cls, dup)
        head[anchors, cols] = cls_p
        prod = head[anchors, cols] * head[anchors, 4]
        # enforce uniqueness after rounding
        for _ in range(8):
            _, first = np.unique(prod, return_index=True)
            clash = np.setdiff1d(np.arange(n_cand), first)
            if clash.size == 0:
                break
            head[anchors[clash], cols[clash]] = np.nextafter(head[anchors[clash], cols[clash]], np.float32(0))
            prod = head[anchors, cols] * head[anchors, 4]
    out = head if anchor_major else np.ascontiguousarray(head.T)
    if return_centers:
        return out, centers
    return out


class DenseScene:
    """Config 5: ``n_objects`` long-lived objects x ``dup`` near-duplicate anchors, boxes moving
    <= 2 px/frame, 10 classes (~1800 candidates -> ~300 kept per frame)."""

    def __init__(self, seed: int, n_objects: int = 300, dup: int = 6, n_obj_classes: int = 10,
                 n_classes_plus4: int = 84, n_anchors: int = 8400, input_hw=(640, 640), orbit: Optional[float] = None,
                 jitter: float = 1.0):
        """``orbit`` (network-input pixels): instead of drifting linearly -- which walks the objects into each other
        after ~40 frames -- every object circles its grid cell with that radius at <= 0.6 px/frame, so a run of any
        length keeps the same ~``n_objects`` separate objects (the long-lived tracks of config 5)."""
        self.seed, self.n_objects, self.dup = seed, n_objects, dup
        self.orbit, self.jitter = orbit, jitter
        self.n_obj_classes, self.c, self.a, self.input_hw = n_obj_classes, n_classes_plus4, n_anchors, input_hw
        rng = np.random.default_rng(seed)
        in_h, in_w = input_hw
        # keep away from the letterbox bars and spread on a jittered grid so most objects survive NMS
        g = int(np.ceil(np.sqrt(n_objects * 16 / 9)))
        gx, gy = np.meshgrid(np.arange(g), np.arange(int(np.ceil(n_objects / g))))
        cells = np.stack([gx.ravel(), gy.ravel()], axis=1)[:n_objects].astype(np.float64)
        ny = cells[:, 1].max() + 1
        self.centers = np.stack([
            (cells[:, 0] + 0.5) / g * in_w * 0.9 + 0.05 * in_w,
            (cells[:, 1] + 0.5) / ny * in_h * 0.5 + 0.25 * in_h,
            rng.uniform(14, 24, n_objects), rng.uniform(10, 18, n_objects)], axis=1)
        self.vel = rng.uniform(-0.6, 0.6, size=(n_objects, 2))
        self.obj_cls = rng.integers(0, n_obj_classes, n_objects)

    def head(self, t: int) -> np.ndarray:
        centers = self.centers.copy()
        if self.orbit is None:
            centers[:, :2] += self.vel * t
        else:
            speed = np.hypot(self.vel[:, 0], self.vel[:, 1])  # <= 0.85 px/frame along the circle
            phase = np.arctan2(self.vel[:, 1], self.vel[:, 0]) + speed / self.orbit * t
            centers[:, 0] += self.orbit * np.cos(phase)
            centers[:, 1] += self.orbit * np.sin(phase)
        # same objects and classes every frame; new anchors, jitter and scores per frame
        return synth_head(self.seed * 1000003 + t, self.c, self.a, self.n_objects, self.dup,
                          self.n_obj_classes, self.input_hw, False, self.jitter, centers, self.obj_cls)

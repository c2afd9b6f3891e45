"""Golden vectors that pin oracle/ultralytics_restate.py where something installed can pin it:
``torchvision.ops.nms`` (the NMS step of ops.non_max_suppression) and torch CPU float32 arithmetic
(the class shift ``boxes + cls * 7680`` and ``scale_boxes``: subtract pad, ``/= gain``, clamp).
Run in the build container:  python tests/golden/make_golden_ultralytics.py  -> tests/golden/ultralytics.npz
(ultralytics itself is not installed: the composition of these steps stays "parity unpinned".)"""
import os
import sys

import numpy as np
import torch
import torchvision

HERE = os.path.dirname(os.path.abspath(__file__))
out = {"versions": np.array([torch.__version__, torchvision.__version__])}
rng = np.random.default_rng(20261018)


def boxes_case(n, span, size, ties):
    cx, cy = rng.uniform(0, span, n), rng.uniform(0, span, n)
    w, h = rng.uniform(4, size, n), rng.uniform(4, size, n)
    b = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1).astype(np.float32)
    s = rng.uniform(0.05, 1.0, n).astype(np.float32)
    if ties:
        s = np.round(s * 16) / np.float32(16)  # many equal scores: exercises the stable sort
    return b, s


cases = [(0, 100, 50, False), (1, 100, 50, False), (40, 200, 120, False), (300, 640, 200, False), (300, 640, 200, True),
         (1500, 640, 90, False), (64, 50, 60, True)]
for k, (n, span, size, ties) in enumerate(cases):
    b, s = boxes_case(n, span, size, ties)
    if n >= 40:  # degenerate boxes: zero area and duplicates (0 / 0 IoU is NaN and must not suppress)
        b[3] = b[2]
        b[5, 2:] = b[5, :2]
        b[7] = (10, 10, 10, 10)
        b[8] = (10, 10, 10, 10)
    for t, thr in enumerate((0.45, 0.5, 0.7)):
        keep = torchvision.ops.nms(torch.from_numpy(b), torch.from_numpy(s), thr).numpy()
        out[f"nms{k}_{t}_keep"] = keep
    out[f"nms{k}_boxes"], out[f"nms{k}_scores"] = b, s
out["nms_n"] = np.array([len(cases)])

# class shift and scale_boxes arithmetic on torch CPU float32 tensors
b, _ = boxes_case(500, 640, 300, False)
cls = rng.integers(0, 80, 500)
c = torch.from_numpy(cls.astype(np.float32))[:, None] * 7680
out["shift_boxes"], out["shift_cls"] = b, cls
out["shift_out"] = (torch.from_numpy(b) + c).numpy()
shapes = [((384, 640), (1080, 1920)), ((640, 640), (1080, 1920)), ((640, 384), (1920, 1080)), ((640, 640), (2160, 3840)),
          ((480, 640), (723, 1001)), ((640, 640), (360, 640))]
for k, (img1, img0) in enumerate(shapes):
    t = torch.from_numpy(b.copy())
    gain = min(img1[0] / img0[0], img1[1] / img0[1])
    pad = (round((img1[1] - img0[1] * gain) / 2 - 0.1), round((img1[0] - img0[0] * gain) / 2 - 0.1))
    t[..., 0] -= pad[0]
    t[..., 1] -= pad[1]
    t[..., 2] -= pad[0]
    t[..., 3] -= pad[1]
    t[..., :4] /= gain
    t[..., 0] = t[..., 0].clamp(0, img0[1])
    t[..., 1] = t[..., 1].clamp(0, img0[0])
    t[..., 2] = t[..., 2].clamp(0, img0[1])
    t[..., 3] = t[..., 3].clamp(0, img0[0])
    out[f"scale{k}_shapes"] = np.array([*img1, *img0])
    out[f"scale{k}_out"] = t.numpy()
out["scale_n"] = np.array([len(shapes)])
np.savez_compressed(os.path.join(HERE, "ultralytics.npz"), **out)
print("wrote", os.path.join(HERE, "ultralytics.npz"), {k: v.shape for k, v in list(out.items())[:6]})

#!/usr/bin/env python
"""Drive the batched engine on synthetic streams (needs a B200).

    python examples/run_engine.py --streams 32 --height 2160 --width 3840 --ticks 20 --roi --motion

Frames come from `synth.MotionScene` (moving rectangles + noise; every 5th stream static so the
motion gate and adaptive FPS kick in), the "detector" is a stand-in that returns a synthetic decoded
YOLOv8 head for each active stream.  Prints per-tick counts the way the reference's metrics sink
receives them (pipeline.py:184-189).
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from realtime_video_analytics_32streams_b200 import (DetectorConfig, HotPathEngine, StreamConfig, TrackerConfig,  # noqa: E402
                                                     synth)

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=8)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--ticks", type=int, default=10)
ap.add_argument("--roi", action="store_true")
ap.add_argument("--motion", action="store_true")
args = ap.parse_args()

H, W = args.height, args.width
streams = [StreamConfig(name=f"cam{i:02d}", target_fps=25,
                        roi_polygons=synth.synth_polygons(4000 + i, H, W) if args.roi else None,
                        motion_filter=args.motion, adaptive_fps=args.motion, min_target_fps=5, idle_frame_tolerance=5)
           for i in range(args.streams)]
scenes = [synth.MotionScene(100 + i, H, W, static=(i % 5 == 4)) for i in range(args.streams)]
heads = [synth.DenseScene(200 + i, n_objects=12, dup=3, n_obj_classes=5) for i in range(args.streams)]
state = {"t": 0, "active": []}


def infer(tensor):
    """Stand-in for the detector forward: one synthetic head per frame that reached the detector."""
    return torch.from_numpy(np.stack([heads[i % len(heads)].head(state["t"]) for i in range(tensor.shape[0])])).cuda()


engine = HotPathEngine(streams, DetectorConfig(confidence_threshold=0.35, iou_threshold=0.5),
                       TrackerConfig(max_age=30, max_iou_distance=0.5, min_hits=1), infer=infer)
for t in range(args.ticks):
    state["t"] = t
    frames = [sc.frame(t) for sc in scenes]
    t0 = time.perf_counter()
    results = engine.tick(frames)
    dt = (time.perf_counter() - t0) * 1e3
    processed = sum(r.processed for r in results)
    print(f"tick {t:3d}: {dt:7.2f} ms  processed {processed}/{len(results)}  "
          f"detections {sum(r.n_detections for r in results)}  tracks {sum(r.n_tracks for r in results)}  "
          f"skips {[r.skip_reason for r in results if not r.processed][:4]}")
engine.h.poll_status()

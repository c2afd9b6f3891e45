"""Writes config/pipeline-4k-roi.yaml: BASELINE.json configs[3] as a file the reference's own loader reads
(32 x 4K streams, per-stream hexagon + triangle ROI from synth.synth_polygons(4000 + i), motion gate, adaptive FPS).

    python tools/make_config4_yaml.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from realtime_video_analytics_32streams_b200 import synth  # noqa: E402

HEAD = """# BASELINE.json configs[3]: 32 streams of 4K (H.265-shape frames) with per-stream ROI polygons and
# motion-adaptive FPS.  The reference ships no YAML that sets any ROI / motion / adaptive key
# (SURVEY.md §5), so this file is authored here (tools/make_config4_yaml.py) using only keys the reference's
# StreamConfig already has (src/realtime_analytics/config.py:57-73); it loads with the reference's own
# load_config once `backend: b200` is registered (INTEGRATION.md §3).  The polygons of stream i are
# synth.synth_polygons(4000 + i, 2160, 3840): the ones tools/bench_configs.py and the tests rasterise.
detector:
  backend: b200
  model_path: models/yolo/yolov8s.pt
  model_type: yolov8
  confidence_threshold: 0.35
  iou_threshold: 0.5
  input_size: [640, 640]
  half: false
tracker:
  type: b200_iou
  max_age: 30
  max_iou_distance: 0.5
  min_hits: 1
max_concurrent_streams: 32
streams:
"""


def main():
    out = [HEAD]
    for i in range(32):
        polys = synth.synth_polygons(4000 + i, 2160, 3840)
        out.append(f"  - name: cam-4k-{i:02d}\n    url: rtsp://camera-{i:02d}/stream\n    target_fps: 25\n    roi_polygons:\n")
        for p in polys:
            out.append("      - [" + ", ".join(f"[{x}, {y}]" for x, y in p) + "]\n")
        out.append(f"    motion_filter: true\n    motion_threshold: {(0.02, 0.01, 0.004)[i % 3]}\n    adaptive_fps: true\n"
                   "    min_target_fps: 5\n    idle_frame_tolerance: 60\n")
    with open(os.path.join(ROOT, "config", "pipeline-4k-roi.yaml"), "w") as fh:
        fh.write("".join(out))


if __name__ == "__main__":
    main()

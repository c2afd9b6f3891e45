"""Per-step CPU time of the reference's own call sequence (oracle, cv2 back end: the OpenCV / NumPy calls
frame_filter.py, detector.py and tracker.py make) on this box's host cores -- SURVEY.md §8(d) "timing of the
reference CPU path".  One thread (cv2.setNumThreads(1)); medians over `--frames` frames after 3 warm-up frames.
Reported baseline only.  Writes JSON to --out."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
import numpy as np
from oracle import hotpath as O
from realtime_video_analytics_32streams_b200 import synth

ap = argparse.ArgumentParser(); ap.add_argument("--frames", type=int, default=12); ap.add_argument("--out", default="")
args = ap.parse_args()
cv2.setNumThreads(1)


def run(name, hw, n_obj, dup, roi, motion):
    H, W = hw
    frames = [synth.synth_frame(100 + t, H, W) for t in range(2)]
    polys = synth.synth_polygons(4000, H, W) if roi else None
    scene = synth.DenseScene(9000, n_objects=n_obj, dup=dup, n_obj_classes=10)
    trk = O.IouTracker(30, 0.5, 1)
    mf = O.MotionFilter(0.02, backend="cv2") if motion else None
    steps = {k: [] for k in ("roi", "motion", "preprocess", "postprocess", "tracker", "total")}
    for t in range(args.frames + 3):
        f = frames[t % 2]
        head = scene.head(t)[None]
        t0 = time.perf_counter()
        if roi:
            f = O.apply_roi(f, polys, backend="cv2")
        t1 = time.perf_counter()
        if mf is not None:
            mf.should_process(f)
        t2 = time.perf_counter()
        tensor, meta = O.preprocess(f, (640, 640), False, backend="cv2")
        t3 = time.perf_counter()
        dets = O.filter_detections(O.postprocess(head, meta, 0.35, 0.5), 0.35)
        t4 = time.perf_counter()
        tracks = trk.update("s", dets)
        t5 = time.perf_counter()
        if t >= 3:
            for k, v in zip(steps, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t5 - t0)):
                steps[k].append(v * 1e3)
    r = {k: round(float(np.median(v)), 3) for k, v in steps.items()}
    r.update(config=name, frame=[H, W], detections=len(dets), tracks=len(tracks), unit="ms per frame, 1 thread",
             frames_per_s_1_thread=round(1e3 / r["total"], 2))
    return r


out = [run("3: 1080p, 72 candidates -> 24 kept (bench.py workload)", (1080, 1920), 24, 3, False, False),
       run("5: 1080p dense, ~1800 candidates -> ~300 kept", (1080, 1920), 300, 6, False, False),
       run("4: 4K + ROI polygons + motion gate", (2160, 3840), 24, 3, True, True)]
res = {"cores_used": 1, "cores_available": len(os.sched_getaffinity(0)), "cv2": cv2.__version__, "numpy": np.__version__,
       "results": out}
for r in out:
    print(json.dumps(r))
if args.out:
    json.dump(res, open(args.out, "w"), indent=1)

"""Head-decode kernel variants on one GPU: device time of the decode phase alone (b200va_set_profiling events
around the decode launch), detections compared against the register kernel's.

    python tools/bench_decode.py [--batch 32] [--steps 40] [--out profiles/r2_decode_sweep.json]

Variants are selected with the library's B200VA_DECODE_* developer knobs (read at b200va_create):
impl 1 = k_decode_cm (registers), 2 = k_decode_ring (TMA-fed persistent ring), 3 = k_decode_cm_split."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from realtime_video_analytics_32streams_b200 import _native, synth

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, nargs="+", default=[32])
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--out", default="")
ap.add_argument("--quick", action="store_true")
ap.add_argument("--only-ring", action="store_true", help="just the automatic ring variant (for ncu)")
args = ap.parse_args()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
C, A, SETS = 84, 8400, 4
dev = torch.device("cuda", 0)


def run(B, env):
    for k in list(os.environ):
        if k.startswith("B200VA_DECODE_"):
            del os.environ[k]
    os.environ.update({k: str(v) for k, v in env.items()})
    h = _native.Handle(device=0, max_batch=B, max_anchors=A, max_candidates=2048, max_dets=512, max_streams=B, max_tracks=64)
    h.set_profiling(True)
    metas = (_native.Letterbox * B)(*[_native.letterbox_meta(1080, 1920, 640, 640) for _ in range(B)])
    dets = h.alloc_dets(B)
    for k in range(5):
        h.postprocess(HEADS[B][k % SETS], metas, 0.35, 0.5, filter_conf=0.35, out=dets)
    h.phase_times()
    ts, tn = [], []
    for k in range(args.steps):
        # a 256 MB fill first: the host runs ahead of the GPU (no launch latency inside the event pair) and the L2 is
        # left full of dirty lines, like after the letterbox of a real tick (the decode's reads then pay their write-back)
        FLUSH.fill_(k & 1)
        h.postprocess(HEADS[B][k % SETS], metas, 0.35, 0.5, filter_conf=0.35, out=dets)
        pt = h.phase_times()
        ts.append(pt["decode"])
        tn.append(pt["nms"])
    h.postprocess(HEADS[B][0], metas, 0.35, 0.5, filter_conf=0.35, out=dets)
    torch.cuda.synchronize()
    sig = {k: v.cpu().numpy().copy() for k, v in dets.items() if not k.startswith("_")}
    h.poll_status()
    h.close()
    ms = float(np.median(ts))
    gbs = B * C * A * 4 / (ms * 1e-3) / 1e9
    return {"env": env, "batch": B, "decode_us": round(ms * 1e3, 2), "decode_us_min": round(min(ts) * 1e3, 2),
            "nms_us": round(float(np.median(tn)) * 1e3, 2), "GBps": round(gbs, 1), "frac_of_measured_peak": round(gbs / PEAK, 3)}, sig


def same(a, b):
    n = a["count"]
    if not np.array_equal(n, b["count"]):
        return False
    for i, k in enumerate(n):
        for key in ("bbox_xyxy", "conf", "cls"):
            if not np.array_equal(a[key][i, :k], b[key][i, :k]):
                return False
    return True


HEADS = {}
FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
results = []
for B in args.batch:
    scenes = [synth.DenseScene(7000 + s, n_objects=24, dup=3, n_obj_classes=10) for s in range(B)]
    HEADS[B] = [torch.from_numpy(np.stack([sc.head(t) for sc in scenes])).to(dev) for t in range(SETS)]
    if args.only_ring:
        r, _ = run(B, {"B200VA_DECODE_IMPL": 2})
        print(json.dumps(r), flush=True)
        continue
    base, ref = run(B, {"B200VA_DECODE_IMPL": 1})
    base["name"] = "k_decode_cm<4> (registers)"
    results.append(base)
    print(json.dumps(base), flush=True)
    variants = [("k_decode_cm_split<8,12>", {"B200VA_DECODE_IMPL": 3}), ("k_decode_ring auto", {"B200VA_DECODE_IMPL": 2})]
    if not args.quick:
        for ta in (128, 192, 224, 256, 384, 512):
            for per_sm in (2, 3, 4):
                variants.append((f"ring ta={ta} ctas/sm={per_sm}", {"B200VA_DECODE_IMPL": 2, "B200VA_DECODE_TA": ta,
                                                                    "B200VA_DECODE_CTAS_PER_SM": per_sm}))
        for rows, stages in ((6, 8), (12, 4), (12, 6), (14, 4), (21, 2), (21, 3), (28, 2), (42, 2)):
            variants.append((f"ring ta=256 rows={rows} stages={stages}", {"B200VA_DECODE_IMPL": 2, "B200VA_DECODE_TA": 256,
                                                                          "B200VA_DECODE_ROWS": rows, "B200VA_DECODE_STAGES": stages}))
    for name, env in variants:
        try:
            r, sig = run(B, env)
        except Exception as exc:  # a variant the library refuses (shared memory, ...) is reported, not fatal
            r, sig = {"env": env, "batch": B, "error": str(exc)}, None
        r["name"] = name
        if sig is not None:
            r["identical_to_register_kernel"] = same(ref, sig)
        results.append(r)
        print(json.dumps(r), flush=True)
if args.out:
    with open(os.path.join(ROOT, args.out), "w") as fh:
        json.dump({"peak_GBps": PEAK, "head": [C, A], "results": results}, fh, indent=1)

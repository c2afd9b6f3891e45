"""Device-time per tick for the BASELINE.json configurations that are not the bench.py line
(configs 1, 2, 4, 5 of SURVEY.md §8d).  Inputs resident in HBM, CUDA events around every C-ABI call,
medians over `--steps` ticks.  Writes profiles/r1_configs.json when --out is given."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from realtime_video_analytics_32streams_b200 import _native, synth

ap = argparse.ArgumentParser(); ap.add_argument("--steps", type=int, default=30); ap.add_argument("--out", default="")
ap.add_argument("--only", default="")
ap.add_argument("--lshape", type=int, default=-1, help="with --only L: run just this letterbox shape (0-5)")
args = ap.parse_args()
dev = torch.device("cuda", 0)
PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0


def timed(fns, steps, warm=5):
    """fns: list of (name, callable). Returns {name: median ms} and total."""
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(len(fns) + 1)] for _ in range(steps)]
    for k in range(warm):
        for _, f in fns: f(k)
    torch.cuda.synchronize()
    for k in range(steps):
        ev[k][0].record()
        for i, (_, f) in enumerate(fns):
            f(warm + k); ev[k][i + 1].record()
    torch.cuda.synchronize()
    res = {n: float(np.median([e[i].elapsed_time(e[i + 1]) for e in ev])) for i, (n, _) in enumerate(fns)}
    res["tick_total"] = float(np.median([e[0].elapsed_time(e[-1]) for e in ev]))
    return res


def simple(name, B, hw, n_obj, dup, conf=0.35, iou=0.5, trk=(30, 1, 0.5)):
    H, W = hw
    h = _native.Handle(device=0, max_batch=max(B, 1), max_anchors=8400, max_candidates=4096, max_dets=1024, max_streams=B, max_tracks=2048)
    g = torch.Generator(device=dev); g.manual_seed(5)
    sets = 4
    frames = [torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev, generator=g) for _ in range(sets)]
    batches = [_native.FrameBatch(list(f.unbind(0))) for f in frames]
    scenes = [synth.DenseScene(9000 + s, n_objects=n_obj, dup=dup, n_obj_classes=10) for s in range(B)]
    heads = [torch.from_numpy(np.stack([sc.head(t) for sc in scenes])).to(dev) for t in range(sets)]
    metas = (_native.Letterbox * B)(*[_native.letterbox_meta(H, W, 640, 640) for _ in range(B)])
    net = torch.empty((B, 3, 640, 640), dtype=torch.float32, device=dev)
    dets, tracks = h.alloc_dets(B), h.alloc_tracks(B)
    slots = _native._int_array(list(range(B)))
    r = timed([("preprocess", lambda k: h.preprocess(batches[k % sets], (640, 640), 0, out=net)),
               ("postprocess", lambda k: h.postprocess(heads[k % sets], metas, conf, iou, filter_conf=conf, out=dets)),
               ("tracker", lambda k: h.tracker_update(slots, dets, trk[0], trk[1], trk[2], out=tracks))], args.steps)
    h.poll_status()
    # the same tick as one prepared b200va_tick replayed from a CUDA graph (what a deployed loop runs: no host
    # launch gaps -- with 1-4 streams three Python calls cost more than the kernels -- and NMS + tracker under the letterbox)
    plans = [h.plan_tick(frames=batches[k], net_out=net, dst_hw=(640, 640), head=heads[k], metas=metas, conf_thr=conf,
                         iou_thr=iou, filter_conf=conf, dets=dets, slots=slots, tracker_cfg=trk, tracks=tracks)
             for k in range(sets)]
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream()); graphs = []
    with torch.cuda.stream(side):
        for k in range(sets):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=side):
                h.tick(plans[k])
            graphs.append(gr)
    torch.cuda.current_stream().wait_stream(side)
    for k in range(5): graphs[k % sets].replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    n_rep = max(args.steps, 50)
    for k in range(n_rep): graphs[k % sets].replay()
    e1.record(); torch.cuda.synchronize()
    r["tick_graph"] = e0.elapsed_time(e1) / n_rep
    h.poll_status()
    r.update(config=name, streams=B, frame=[H, W], frames_per_s=B / (r["tick_graph"] * 1e-3),
             frames_per_s_serial_calls=B / (r["tick_total"] * 1e-3),
             dets_per_frame=float(dets["count"].float().mean()), tracks_per_stream=float(tracks["count"].float().mean()))
    h.close()
    return r


def config4(B=32):
    """32 x 4K with per-stream ROI polygons and the motion gate (every pixel is read)."""
    H, W = 2160, 3840
    h = _native.Handle(device=0, max_batch=B, max_anchors=8400, max_candidates=4096, max_dets=1024, max_streams=B, max_tracks=2048)
    g = torch.Generator(device=dev); g.manual_seed(7)
    sets = 2
    frames = [torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev, generator=g) for _ in range(sets)]
    masks = [h.roi_rasterize(synth.synth_polygons(4000 + s, H, W), H, W) for s in range(B)]
    batches = [_native.FrameBatch(list(f.unbind(0)), masks) for f in frames]
    gray = [[torch.empty((H, W), dtype=torch.uint8, device=dev) for _ in range(B)] for _ in range(2)]
    changed = torch.empty((B,), dtype=torch.int32, device=dev)
    h.motion(batches[0], [None] * B, gray[1], changed_out=changed)
    scenes = [synth.DenseScene(9100 + s, n_objects=24, dup=3, n_obj_classes=10) for s in range(B)]
    heads = [torch.from_numpy(np.stack([sc.head(t) for sc in scenes])).to(dev) for t in range(sets)]
    metas = (_native.Letterbox * B)(*[_native.letterbox_meta(H, W, 640, 640) for _ in range(B)])
    net = torch.empty((B, 3, 640, 640), dtype=torch.float32, device=dev)
    dets, tracks = h.alloc_dets(B), h.alloc_tracks(B)
    slots = _native._int_array(list(range(B)))
    r = timed([("motion(+roi)", lambda k: h.motion(batches[k % sets], gray[(k + 1) % 2], gray[k % 2], changed_out=changed)),
               ("preprocess(+roi)", lambda k: h.preprocess(batches[k % sets], (640, 640), 0, out=net)),
               ("postprocess", lambda k: h.postprocess(heads[k % sets], metas, 0.35, 0.5, filter_conf=0.35, out=dets)),
               ("tracker", lambda k: h.tracker_update(slots, dets, 30, 1, 0.5, out=tracks))], args.steps)
    h.poll_status()
    # the same two steps as ONE pass over each frame (b200va_motion_preprocess)
    rf = timed([("fused", lambda k: h.motion_preprocess(batches[k % sets], gray[(k + 1) % 2], gray[k % 2], (640, 640), 0,
                                                        changed_out=changed, out=net))], args.steps)
    rp = timed([("fused", lambda k: h.motion_preprocess(batches[k % sets], gray[(k + 1) % 2], gray[k % 2], (640, 640),
                                                        0 | _native.OUT_FLAG_PADS_VALID, changed_out=changed, out=net))],
               args.steps)
    h.poll_status()
    fused_bytes = B * (H * W * 3 + 3 * H * W + 3 * 640 * 640 * 4)  # frame + mask + prev gray + new gray + network input
    r["motion+preprocess fused"] = rf["fused"]
    r["motion+preprocess fused, pad rows kept"] = rp["fused"]
    r["fused_GBps"] = fused_bytes / (rf["fused"] * 1e-3) / 1e9
    r["fused_frac_of_peak"] = r["fused_GBps"] / PEAK
    r["fused_algorithmic_bytes"] = fused_bytes
    r["tick_total_fused"] = rf["fused"] + r["postprocess"] + r["tracker"]
    motion_bytes = B * (H * W * 3 + 3 * H * W)  # frame + mask + prev gray + new gray
    pre_bytes = B * (720 * W * 3 + 720 * W + 3 * 640 * 640 * 4)  # tapped rows (+ their mask rows) + output
    r.update(config="4: 32x4K + ROI + motion", streams=B, frame=[H, W], frames_per_s=B / (r["tick_total_fused"] * 1e-3),
             frames_per_s_separate_kernels=B / (r["tick_total"] * 1e-3),
             motion_GBps=motion_bytes / (r["motion(+roi)"] * 1e-3) / 1e9, motion_frac_of_peak=motion_bytes / (r["motion(+roi)"] * 1e-3) / 1e9 / PEAK,
             preprocess_GBps=pre_bytes / (r["preprocess(+roi)"] * 1e-3) / 1e9,
             preprocess_frac_of_peak=pre_bytes / (r["preprocess(+roi)"] * 1e-3) / 1e9 / PEAK,
             motion_algorithmic_bytes=motion_bytes, preprocess_algorithmic_bytes=pre_bytes)
    h.close()
    return r


def letterbox_only(name, B, hw, fmt=0, mask=False):
    H, W = hw
    h = _native.Handle(device=0, max_batch=B, max_anchors=8400, max_candidates=1024, max_dets=256, max_streams=B, max_tracks=256)
    g = torch.Generator(device=dev); g.manual_seed(5)
    frames = [torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev, generator=g) for _ in range(2)]
    masks = [h.roi_rasterize(synth.synth_polygons(4000 + s, H, W), H, W) for s in range(B)] if mask else None
    batches = [_native.FrameBatch(list(f.unbind(0)), masks) for f in frames]
    dt = {0: torch.float32, 1: torch.float16}.get(fmt, torch.uint8)
    net = torch.empty((B, 3, 640, 640), dtype=dt, device=dev)
    r = timed([("preprocess", lambda k: h.preprocess(batches[k % 2], (640, 640), fmt, out=net))], args.steps)
    m = _native.letterbox_meta(H, W, 640, 640)
    ys = set()
    from oracle import cv_restate as cvr
    y0, y1, b0, b1 = cvr.linear_taps(H, m.new_h, False)
    rows = len(set(y0[b0 != 0].tolist()) | set(y1[b1 != 0].tolist()))
    esz = {0: 4, 1: 2}.get(fmt, 1)
    alg = B * (rows * W * 3 + (rows * W if mask else 0) + 3 * 640 * 640 * esz)
    r.update(config=name, streams=B, frame=[H, W], algorithmic_bytes=alg, GBps=alg / (r["preprocess"] * 1e-3) / 1e9,
             frac_of_peak=alg / (r["preprocess"] * 1e-3) / 1e9 / PEAK, tapped_rows=rows)
    h.close()
    return r


def dfl_only(B=32, nc=80):
    """a14: raw Detect head [B, 64 + nc, 8400] -> decoded [B, 4 + nc, 8400]."""
    h = _native.Handle(device=0, max_batch=B, max_anchors=8400, max_candidates=1024, max_dets=256, max_streams=B, max_tracks=256)
    g = torch.Generator(device=dev); g.manual_seed(9)
    raws = [torch.randn((B, 64 + nc, 8400), dtype=torch.float32, device=dev, generator=g) * 3 for _ in range(3)]
    out_t = torch.empty((B, 4 + nc, 8400), dtype=torch.float32, device=dev)
    r = timed([("dfl_decode", lambda k: h.dfl_decode(raws[k % 3], nc, out=out_t))], args.steps)
    alg = B * 8400 * 4 * ((64 + nc) + (4 + nc))
    r.update(config="a14: DFL decode of 32 raw heads [144, 8400]", streams=B, algorithmic_bytes=alg,
             GBps=alg / (r["dfl_decode"] * 1e-3) / 1e9, frac_of_peak=alg / (r["dfl_decode"] * 1e-3) / 1e9 / PEAK)
    h.close()
    return r


out = []
todo = args.only.split(",") if args.only else ["1", "2", "5", "4", "D"]
if "L" in todo:
    shapes = [("letterbox 32x1080p fp32", 32, (1080, 1920), 0, False), ("letterbox 32x1080p fp16", 32, (1080, 1920), 1, False),
              ("letterbox 32x4K fp32", 32, (2160, 3840), 0, False), ("letterbox 32x4K fp32 + ROI mask", 32, (2160, 3840), 0, True),
              ("letterbox 32x720p fp32", 32, (720, 1280), 0, False),
              ("letterbox 32x1440p fp32 (non-integer ratio)", 32, (1440, 2560), 0, False)]
    for i, (nm, b, hw, fmt, mask) in enumerate(shapes):
        if args.lshape < 0 or args.lshape == i:
            out.append(letterbox_only(nm, b, hw, fmt=fmt, mask=mask))
if "1" in todo: out.append(simple("1: 1 stream 1080p (pipeline-sim shape)", 1, (1080, 1920), 10, 1))
if "2" in todo: out.append(simple("2: 4 streams 1080p (pipeline-rtsp shape)", 4, (1080, 1920), 10, 1))
if "5" in todo: out.append(simple("5: dense stress, 32 streams, ~1800 candidates -> ~300 kept", 32, (1080, 1920), 300, 6))
if "4" in todo: out.append(config4())
if "D" in todo: out.append(dfl_only())
for r in out:
    print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()}))
if args.out:
    json.dump({"peak_GBps": PEAK, "results": out}, open(args.out, "w"), indent=1)

#!/usr/bin/env python
"""Benchmark of the pre + post + track hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--schedule 0|1|2] [--no-graph] [--no-cpu]

Workload at every N: each GPU serves 32 streams of 1080p BGR frames with a synthetic decoded
YOLOv8 head [32, 84, 8400] (config 3 of BASELINE.json, the one the metric is quoted on); with
N > 1 every rank (one process per GPU, torchrun) serves its own 32 streams -- weak scaling, no
data-path collective.  One "step" is one tick: letterbox preprocess of the 32 frames, head decode +
NMS of the 32 heads, tracker update of the 32 streams -- one prepared `b200va_tick` call (letterbox on
the caller's stream, decode -> NMS -> tracker on the library's second stream, fork and join inside the
call).  The detector forward is outside the measured path (north_star): the head tensors stand in
for its output.

`value`   : frames/s with frames and heads already resident in HBM (CUDA events around the K steps,
            max over ranks).  The tick is replayed from a CUDA graph per input set; every 10th step is
            launched eagerly with an event pair around the letterbox launch (the live kernel timing of
            `roofline`).  Four rotating input sets keep every step's reads out of the 126 MB L2.
`roofline`: the letterbox kernel's algorithmic bytes / that event-timed duration vs the measured HBM peak.
`e2e`     : the same tick through the public API (HotPathEngine.submit / collect, two ticks in flight)
            with HOST buffers: pinned frames and heads are copied to the device and the result tables
            are read back inside the timed region.
`cpu_baseline` / `--impl reference`: the reference's own OpenCV/NumPy call sequence (oracle, cv2
            back end) on the host cores of the same box (one thread / one process per core).
Informational keys: value_eager_tick, value_three_serial_calls (the unfused call sequence),
value_pad_rows_written_once (B200VA_OUT_FLAG_PADS_VALID), value_32_streams_total_strong_scaling (N > 1).
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from realtime_video_analytics_32streams_b200 import synth  # noqa: E402

METRIC = "pre+post+track frames/sec (32x1080p)"
STREAMS = 32
H, W = 1080, 1920
IN_HW = (640, 640)
C, A = 84, 8400
CONF, IOU = 0.35, 0.5
TRK = dict(max_age=30, max_iou_distance=0.5, min_hits=1)
N_SETS = 4  # rotating input sets: every step reads frames / heads that left L2 long ago
OBJECTS, DUP = 24, 3
# algorithmic HBM bytes of the letterbox kernel per 1080p frame: the 360 tapped
# source rows (every third row carries weight 2048, the rest 0) + the fp32 NCHW output
LETTERBOX_BYTES_PER_FRAME = 360 * 1920 * 3 + 3 * 640 * 640 * 4


def workload_config(n_gpus: int) -> dict:
    return {"workload": "32 streams x 1080p BGR per GPU -> letterbox 640x640 fp32 + YOLOv8 head [32,84,8400] decode/NMS "
                        "+ IoU tracker (BASELINE.json configs[2] shape, 32 streams on every GPU)",
            "streams_per_gpu": STREAMS, "streams_total": STREAMS * n_gpus, "frame": [H, W, 3], "head": [C, A],
            "objects_per_frame": OBJECTS * DUP, "conf_thr": CONF, "iou_thr": IOU, "tracker": TRK,
            "cache": f"{N_SETS} rotating input sets (199 MB frames + 90 MB heads each) > 126 MB L2",
            "sharding": "by stream id, no collective"}


def make_heads(stream: int, n_sets: int) -> np.ndarray:
    scene = synth.DenseScene(7000 + stream, n_objects=OBJECTS, dup=DUP, n_obj_classes=10)
    return np.stack([scene.head(t) for t in range(n_sets)])


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, device_index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            # honour CUDA_VISIBLE_DEVICES when mapping the torch ordinal to an NVML index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = device_index
            if vis:
                try:
                    idx = int(vis.split(",")[device_index])
                except (ValueError, IndexError):
                    idx = device_index
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
        except Exception:  # pragma: no cover
            self.nv = None

    def _poll(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.002)

    def start(self):
        if self.nv is not None:
            self._stop.clear()
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join()
            self._thread = None

    def summary(self) -> dict:
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------
# CPU arm: the reference's OpenCV / NumPy call sequence (oracle, cv2 back end)
# --------------------------------------------------------------------------------------------
_W = {}


def _cpu_worker_init(stream_ids, n_sets):
    import cv2

    from oracle import hotpath as O

    cv2.setNumThreads(1)
    _W["O"] = O
    _W["frames"] = {s: synth.synth_frame(3000 + s, H, W) for s in stream_ids}
    _W["heads"] = {s: make_heads(s, n_sets) for s in stream_ids}
    _W["tracker"] = O.IouTracker(TRK["max_age"], TRK["max_iou_distance"], TRK["min_hits"])


def _cpu_tick(args):
    stream_ids, t = args
    O = _W["O"]
    n_tracks = 0
    for s in stream_ids:
        tensor, meta = O.preprocess(_W["frames"][s], IN_HW, False, backend="cv2")
        heads = _W["heads"][s]
        dets = O.postprocess(heads[t % len(heads)][None], meta, CONF, IOU)
        dets = O.filter_detections(dets, CONF)
        n_tracks += len(_W["tracker"].update(f"s{s}", dets))
    return n_tracks


def cpu_single_core_sample(budget_s: float = 10.0, n_streams: int = 4) -> dict:
    """Bounded single-thread sample of the same workload (rank 0, N=1 only)."""
    ids = list(range(n_streams))
    _cpu_worker_init(ids, N_SETS)
    _cpu_tick((ids, 0))  # warm-up
    t0 = time.perf_counter()
    ticks = 0
    while True:
        _cpu_tick((ids, ticks + 1))
        ticks += 1
        if time.perf_counter() - t0 > budget_s or ticks >= 2000:
            break
    dt = time.perf_counter() - t0
    return {"value": round(ticks * n_streams / dt, 2), "unit": "frames/s", "cores": 1, "kind": "port",
            "sample": f"{ticks} ticks x {n_streams} streams of the bench workload (1080p letterbox + [84,8400] decode/NMS + "
                      f"tracker), oracle with the cv2 back end (the reference's own OpenCV/NumPy call sequence), "
                      f"1 thread, {dt:.1f} s"}


def run_reference(args) -> None:
    """--impl reference: the CPU path on every host core (rank 0 only under torchrun)."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import multiprocessing as mp

    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    procs = max(1, min(cores, STREAMS))
    # a step = one tick of a bounded number of streams, spread evenly over the worker processes
    streams_per_proc = 1 if procs >= 8 else 2
    groups = [[p * streams_per_proc + k for k in range(streams_per_proc)] for p in range(procs)]
    n_frames = procs * streams_per_proc
    ctx = mp.get_context("fork")
    pools = [ctx.Pool(1, initializer=_cpu_worker_init, initargs=(g, N_SETS)) for g in groups]

    def tick(t):
        res = [pool.apply_async(_cpu_tick, ((g, t),)) for pool, g in zip(pools, groups)]
        for r in res:
            r.get()

    for t in range(args.warmup):
        tick(t)
    t0 = time.perf_counter()
    for t in range(args.steps):
        tick(args.warmup + t)
    dt = time.perf_counter() - t0
    for pool in pools:
        pool.terminate()
    value = n_frames * args.steps / dt
    sample = (f"each step = one tick of {n_frames} of the {STREAMS} streams ({streams_per_proc} per process, {procs} processes, "
              f"1 OpenCV thread each); same frames/heads/thresholds as the GPU arm")
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 2), "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * dt / args.steps, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": round(value, 2), "unit": "frames/s", "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": round(value, 2), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def run_gpu(args) -> None:
    import torch
    import torch.distributed as dist

    from realtime_video_analytics_32streams_b200 import (DetectorConfig, HotPathEngine, StreamConfig, TrackerConfig,
                                                         _native)

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cores = None
    if world > 1 and not args.no_bind:
        from realtime_video_analytics_32streams_b200.runtime import bind_process_to_gpu

        numa_cores = bind_process_to_gpu(local)  # before any pinned allocation: first touch on the GPU's NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    h = _native.Handle(device=local, max_batch=STREAMS, max_anchors=A, max_candidates=2048, max_dets=512,
                       max_streams=2 * STREAMS, max_tracks=1024)
    # ---- synthetic inputs, resident in HBM --------------------------------------------------
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    frame_sets = [torch.randint(0, 256, (STREAMS, H, W, 3), dtype=torch.uint8, device=dev, generator=g)
                  for _ in range(N_SETS)]
    heads_np = np.stack([make_heads(rank * STREAMS + s, N_SETS) for s in range(STREAMS)], axis=1)  # [sets, 32, C, A]
    head_sets = [torch.from_numpy(heads_np[k]).to(dev) for k in range(N_SETS)]
    metas = (_native.Letterbox * STREAMS)(*[_native.letterbox_meta(H, W, *IN_HW) for _ in range(STREAMS)])
    net_in = torch.empty((STREAMS, 3, *IN_HW), dtype=torch.float32, device=dev)
    dets = h.alloc_dets(STREAMS)
    tracks = h.alloc_tracks(STREAMS)
    slots = _native._int_array(list(range(STREAMS)))
    # argument arrays are built once per input set: the per-step host work is three foreign calls
    batches = [_native.FrameBatch(list(fs.unbind(0))) for fs in frame_sets]

    # one prepared b200va_tick per input set: letterbox on the caller's stream, decode -> NMS -> tracker on the
    # library's second stream (schedule 1: the letterbox starts when the decode kernel is done and overlaps
    # NMS + tracker, so the two HBM-bound kernels never share the bus); fork and join are inside every step
    plans = [h.plan_tick(frames=batches[k], net_out=net_in, dst_hw=IN_HW, fmt=_native.OUT_F32_RGB_NCHW,
                         head=head_sets[k], metas=metas, conf_thr=CONF, iou_thr=IOU, filter_conf=CONF, dets=dets,
                         slots=slots, tracker_cfg=(TRK["max_age"], TRK["min_hits"], TRK["max_iou_distance"]),
                         tracks=tracks, schedule=args.schedule) for k in range(N_SETS)]

    def step(k, ev=None):
        plan = plans[k % N_SETS]
        plan.set_events(*(ev if ev is not None else (None, None)))
        h.tick(plan)

    def serial_step(k):
        h.preprocess(batches[k % N_SETS], IN_HW, _native.OUT_F32_RGB_NCHW, out=net_in)
        h.postprocess(head_sets[k % N_SETS], metas, CONF, IOU, filter_conf=CONF, out=dets)
        h.tracker_update(slots, dets, TRK["max_age"], TRK["min_hits"], TRK["max_iou_distance"], out=tracks)

    # the tick is captured once per input set in a CUDA graph (fork / join included): the timed loop replays
    # graphs, except that every SAMPLE_EVERY-th step runs the same tick eagerly with an event pair around the
    # letterbox launch -- the live kernel timing the roofline is computed from
    SAMPLE_EVERY = 10
    for k in range(max(args.warmup, 3)):
        step(k)
    barrier()
    h.poll_status()
    graphs, kernels_per_graph = None, 0
    if not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            graphs = []
            l0 = h.launch_count
            with torch.cuda.stream(side):
                for k in range(N_SETS):
                    gr = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gr, stream=side):
                        plans[k].set_events(None, None)
                        h.tick(plans[k])
                    graphs.append(gr)
            torch.cuda.current_stream().wait_stream(side)
            kernels_per_graph = (h.launch_count - l0) // N_SETS
            for k in range(max(args.warmup, 3)):
                graphs[k % N_SETS].replay()
        except Exception as exc:  # pragma: no cover - capture refused: every step is launched eagerly instead
            print(f"bench: CUDA graph capture failed ({type(exc).__name__}: {exc}); running eagerly", file=sys.stderr)
            graphs, kernels_per_graph = None, 0
            torch.cuda.synchronize()
        barrier()

    def timed_step(k, ev):
        if graphs is None or ev is not None:
            step(k, ev)
            return 0
        graphs[k % N_SETS].replay()
        return kernels_per_graph

    clocks = ClockSampler(local)
    n_samples = (args.steps + SAMPLE_EVERY - 1) // SAMPLE_EVERY if graphs is not None else args.steps
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_samples)]
    for a, b in kev:  # torch creates the CUDA event on its first record; the library records it afterwards
        a.record()
        b.record()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = h.launch_count
    replayed = 0
    barrier()
    clocks.start()
    start.record()
    for k in range(args.steps):
        sample = graphs is None or k % SAMPLE_EVERY == 0
        replayed += timed_step(args.warmup + k, kev[k // SAMPLE_EVERY if graphs is not None else k] if sample else None)
    end.record()
    barrier()
    clocks.stop()
    launches = h.launch_count - launches0 + replayed
    ms = start.elapsed_time(end)
    k1_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    # per-call device time (separate untimed loop, events around each C-ABI call)
    bev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(20)]
    for k, e in enumerate(bev):
        e[0].record()
        h.preprocess(batches[k % N_SETS], IN_HW, _native.OUT_F32_RGB_NCHW, out=net_in)
        e[1].record()
        h.postprocess(head_sets[k % N_SETS], metas, CONF, IOU, filter_conf=CONF, out=dets)
        e[2].record()
        h.tracker_update(slots, dets, TRK["max_age"], TRK["min_hits"], TRK["max_iou_distance"], out=tracks)
        e[3].record()
    torch.cuda.synchronize()
    breakdown = {name: round(float(np.median([e[i].elapsed_time(e[i + 1]) for e in bev])), 4)
                 for i, name in enumerate(("preprocess", "postprocess(decode+sort_nms)", "tracker_update"))}
    n_tracks = int(tracks["count"].sum().item())
    h.poll_status()

    def timed_variant(fn):
        for k in range(8):
            fn(k)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for k in range(args.steps):
            fn(k)
        s1.record()
        barrier()
        v_ms = s0.elapsed_time(s1)
        if world > 1:
            tmax = torch.tensor([v_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            v_ms = float(tmax.item())
        return world * STREAMS * args.steps / (v_ms * 1e-3)

    # informational: the same tick launched eagerly every step, and as three separate C-ABI calls on one stream
    eager_value = timed_variant(lambda k: step(k))
    serial_value = timed_variant(serial_step)
    # informational: the same tick with B200VA_OUT_FLAG_PADS_VALID -- the 280 pad rows of every 640 x 640 input (44 % of
    # the letterbox output) were written by the earlier steps into the same persistent buffer and are not written
    # again.  Not the headline: `value` rewrites the whole tensor every step, like the reference does.
    pads_value = None
    if graphs is not None:
        p_plans = [h.plan_tick(frames=batches[k], net_out=net_in, dst_hw=IN_HW,
                               fmt=_native.OUT_F32_RGB_NCHW | _native.OUT_FLAG_PADS_VALID, head=head_sets[k], metas=metas,
                               conf_thr=CONF, iou_thr=IOU, filter_conf=CONF, dets=dets, slots=slots,
                               tracker_cfg=(TRK["max_age"], TRK["min_hits"], TRK["max_iou_distance"]), tracks=tracks,
                               schedule=args.schedule) for k in range(N_SETS)]
        side3 = torch.cuda.Stream()
        side3.wait_stream(torch.cuda.current_stream())
        p_graphs = []
        with torch.cuda.stream(side3):
            for k in range(N_SETS):
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=side3):
                    h.tick(p_plans[k])
                p_graphs.append(gr)
        torch.cuda.current_stream().wait_stream(side3)
        pads_value = timed_variant(lambda k: p_graphs[k % N_SETS].replay())
        h.poll_status()
    # informational, N > 1: BASELINE.json's deployment shape -- 32 streams IN TOTAL, 32 / N per GPU (strong scaling).
    # The kernels are latency-bound at 4 streams per launch, so this is far from N x the single-GPU number.
    strong_value = None
    if world > 1 and STREAMS % world == 0 and graphs is not None:
        m = STREAMS // world
        s_dets, s_tracks = h.alloc_dets(m), h.alloc_tracks(m)
        s_slots = _native._int_array([STREAMS + i for i in range(m)])
        s_metas = (_native.Letterbox * m)(*[metas[i] for i in range(m)])
        s_plans = [h.plan_tick(frames=_native.FrameBatch(list(frame_sets[k][:m].unbind(0))), net_out=net_in[:m], dst_hw=IN_HW,
                               fmt=_native.OUT_F32_RGB_NCHW, head=head_sets[k][:m], metas=s_metas, conf_thr=CONF,
                               iou_thr=IOU, filter_conf=CONF, dets=s_dets, slots=s_slots,
                               tracker_cfg=(TRK["max_age"], TRK["min_hits"], TRK["max_iou_distance"]), tracks=s_tracks,
                               schedule=args.schedule) for k in range(N_SETS)]
        side2 = torch.cuda.Stream()
        side2.wait_stream(torch.cuda.current_stream())
        s_graphs = []
        with torch.cuda.stream(side2):
            for k in range(N_SETS):
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=side2):
                    h.tick(s_plans[k])
                s_graphs.append(gr)
        torch.cuda.current_stream().wait_stream(side2)
        strong_value = timed_variant(lambda k: s_graphs[k % N_SETS].replay()) / world  # m * world = 32 frames per step
        h.poll_status()
    if world > 1:
        tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    value = world * STREAMS * args.steps / (ms * 1e-3)

    # ---- end to end through the public API with host buffers ---------------------------------
    streams = [StreamConfig(name=f"r{rank}s{s}") for s in range(STREAMS)]
    host_frames = [torch.empty((H, W, 3), dtype=torch.uint8).pin_memory() for _ in range(STREAMS)]
    for s, hf in enumerate(host_frames):
        hf.copy_(frame_sets[0][s])
    host_heads = [torch.from_numpy(heads_np[k]).pin_memory() for k in range(N_SETS)]
    tick_no = [0]
    heads_from_host = [True]

    def infer(tensor):  # the detector forward is out of scope: its output arrives from pinned host memory
        k = tick_no[0] % N_SETS
        return host_heads[k].to(dev, non_blocking=True) if heads_from_host[0] else head_sets[k]

    eng = HotPathEngine(streams, DetectorConfig(confidence_threshold=CONF, iou_threshold=IOU),
                        TrackerConfig(**TRK), infer=infer, handle=h, input_hw=IN_HW, depth=2)
    eng.tracker._slots = {st.name: STREAMS + s for s, st in enumerate(streams)}
    e2e_steps = max(3, min(args.steps, 100))

    def e2e_run(objects: bool):
        """`e2e_steps` ticks, tick k+1 submitted before tick k is collected (uploads overlap the
        host-side handling of results).  Every tick's H2D and D2H copies are inside the timed region."""
        for s in range(STREAMS):
            h.tracker_reset(STREAMS + s)
        checksum = 0
        for k in range(3):
            tick_no[0] = k
            eng.tick(host_frames)
        barrier()
        moved0 = eng.stager.bytes_moved
        t0 = time.perf_counter()
        prev = None
        for k in range(e2e_steps):
            tick_no[0] = 3 + k
            cur = eng.submit(host_frames)
            if prev is not None:
                for r in eng.collect(prev):
                    checksum += (len(r.tracks) + len(r.detections)) if objects else (r.n_tracks + r.n_detections)
            prev = cur
        for r in eng.collect(prev):
            checksum += (len(r.tracks) + len(r.detections)) if objects else (r.n_tracks + r.n_detections)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            tmax = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dt = float(tmax.item())
        assert checksum > 0
        return world * STREAMS * e2e_steps / dt, dt, (eng.stager.bytes_moved - moved0) // e2e_steps

    e2e_value, e2e_s, frame_bytes_per_step = e2e_run(objects=False)
    e2e_objects, _, _ = e2e_run(objects=True)
    heads_from_host[0] = False
    e2e_dev_heads, _, _ = e2e_run(objects=False)
    ctx0 = eng._ctxs[0]
    d2h = int(ctx0.dets["_flat"].numel() + ctx0.tracks["_flat"].numel())
    h.poll_status()

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            with open(peaks_path) as fh:
                peak, peak_src = float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy, burst)"
        else:
            peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        achieved = LETTERBOX_BYTES_PER_FRAME * STREAMS / (k1_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as fh:
                traffic = json.load(fh).get("k_letterbox_dram_bytes_per_launch")
        line = {"metric": METRIC, "value": round(value, 1), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8/f32", "data": "synthetic",
                "config": workload_config(world),
                "roofline": {"bound": "hbm", "kernel": "k_letterbox<F32_RGB_NCHW> (32 x 1080p per launch)",
                             "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                             "traffic": traffic, "kernel_ms": round(k1_ms, 4),
                             "algorithmic_bytes_per_launch": LETTERBOX_BYTES_PER_FRAME * STREAMS, "peak_source": peak_src},
                "e2e": {"value": round(e2e_value, 1), "unit": "frames/s",
                        "h2d_bytes_per_step": int(frame_bytes_per_step + STREAMS * C * A * 4), "d2h_bytes_per_step": d2h,
                        "h2d_frame_bytes_per_step": int(frame_bytes_per_step), "h2d_head_bytes_per_step": STREAMS * C * A * 4,
                        "steps": e2e_steps, "ms_per_step": round(1e3 * e2e_s / e2e_steps, 3),
                        "value_with_python_objects": round(e2e_objects, 1),
                        "value_heads_resident_on_device": round(e2e_dev_heads, 1),
                        "api": "HotPathEngine.submit/collect (= tick, two ticks in flight) with pinned host frames -> "
                               "FrameResult host arrays (counts, boxes, ids); only the frame rows the letterbox reads "
                               "are uploaded (1 in 3 at 1080p); in `value` the head tensors are ALSO copied from "
                               "pinned host memory every step (conservative: a GPU-resident detector would leave "
                               "them on the device, see value_heads_resident_on_device); value_with_python_objects "
                               "additionally builds every Detection / Track object"},
                "gpu_launches": int(launches), "launches_per_step": launches / max(args.steps, 1),
                "breakdown_ms": breakdown,
                "schedule": {0: "serial", 1: "letterbox after decode, overlapping NMS + tracker (b200va_tick)",
                             2: "letterbox overlapping decode + NMS + tracker (b200va_tick)",
                             3: "letterbox launched beside the decode kernel (programmatic dependent launch), NMS + tracker "
                                "on the second stream (b200va_tick)",
                             4: "software-pipelined b200va_tick: decode + letterbox of step k beside NMS + tracker of step k-1"}[args.schedule],
                "launch": ("CUDA graph replay of the prepared b200va_tick (one graph per input set); every "
                           f"{SAMPLE_EVERY}th step is launched eagerly with an event pair around the letterbox kernel"
                           if graphs is not None else "eager b200va_tick every step, event pair around the letterbox kernel"),
                "letterbox_samples": len(kev),
                "value_eager_tick": round(eager_value, 1), "value_three_serial_calls": round(serial_value, 1),
                "value_pad_rows_written_once": round(pads_value, 1) if pads_value is not None else None,
                "value_32_streams_total_strong_scaling": round(strong_value, 1) if strong_value is not None else None,
                "clocks": clocks.summary(), "tracks_alive": n_tracks,
                "cpu_affinity": (f"rank 0 pinned to {len(numa_cores)} NUMA-local cores" if numa_cores else "unchanged")}
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_single_core_sample()
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--schedule", type=int, default=1, choices=[0, 1, 2, 3, 4],
                    help="b200va_tick schedule: 0 serial, 1 letterbox after decode (default), 2 fully parallel")
    ap.add_argument("--no-graph", action="store_true", help="launch every step eagerly instead of replaying CUDA graphs")
    ap.add_argument("--no-bind", action="store_true", help="N > 1: do not pin each rank to its GPU's NUMA-local cores")
    ap.add_argument("--no-cpu", action="store_true", help="skip the single-core CPU sample")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

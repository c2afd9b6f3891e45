"""Where the kernels of one tick really run (timing build: lib/libb200va_timing.so, `make -C csrc TIMING=1`).
Every kernel stamps %globaltimer at its first CTA start and last CTA end; one graph-replayed tick of the bench
workload is printed per schedule as a timeline relative to the decode kernel's start."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["B200VA_LIB"] = os.path.join(ROOT, "realtime_video_analytics_32streams_b200", "lib", "libb200va_timing.so")
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench as B
from realtime_video_analytics_32streams_b200 import _native

S = int(os.environ.get("STREAMS", B.STREAMS))
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(1)
SETS = 3
frames = [torch.randint(0, 256, (S, B.H, B.W, 3), dtype=torch.uint8, device=dev, generator=g) for _ in range(SETS)]
heads_np = np.stack([B.make_heads(s, SETS) for s in range(S)], axis=1)
heads = [torch.from_numpy(heads_np[k]).to(dev) for k in range(SETS)]
metas = (_native.Letterbox * S)(*[_native.letterbox_meta(B.H, B.W, *B.IN_HW) for _ in range(S)])
out = {}
for schedule in tuple(int(x) for x in os.environ.get("SCHEDULES", "1,3,4").split(",")):
    h = _native.Handle(device=0, max_batch=S, max_anchors=B.A, max_candidates=2048, max_dets=512, max_streams=S,
                       max_tracks=int(os.environ.get("MAX_TRACKS", 1024)))
    h.lib.b200va_debug_read.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_int]
    h.lib.b200va_debug_reset.argtypes = [C.c_void_p]
    nets = [torch.empty((S, 3, *B.IN_HW), dtype=torch.float32, device=dev) for _ in range(SETS)]
    dets, tracks = h.alloc_dets(S), h.alloc_tracks(S)
    batches = [_native.FrameBatch(list(f.unbind(0))) for f in frames]
    plans = [h.plan_tick(frames=batches[k], net_out=nets[k], dst_hw=B.IN_HW, head=heads[k], metas=metas, conf_thr=B.CONF,
                         iou_thr=B.IOU, filter_conf=B.CONF, dets=dets, slots=list(range(S)),
                         tracker_cfg=(30, 1, 0.5), tracks=tracks, schedule=schedule) for k in range(SETS)]
    for k in range(4):
        h.tick(plans[k % SETS])
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graphs = []
    with torch.cuda.stream(side):
        for k in range(SETS):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=side):
                h.tick(plans[k])
            graphs.append(gr)
    torch.cuda.current_stream().wait_stream(side)
    rows = []
    for rep in range(5):
        for k in range(6):
            graphs[k % SETS].replay()
        torch.cuda.synchronize()
        h.lib.b200va_debug_reset(h._h)
        # three back-to-back ticks: stamps accumulate min start / max end over them, so measure ONE tick between syncs
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        graphs[0].replay()  # keeps the GPU busy so the measured tick is not launch-latency bound
        torch.cuda.synchronize()
        h.lib.b200va_debug_reset(h._h)
        e0.record()
        graphs[1].replay()
        e1.record()
        torch.cuda.synchronize()
        buf = (C.c_int64 * 64)()
        h.lib.b200va_debug_read(h._h, buf, 64)
        v = [int(x) & 0xFFFFFFFFFFFFFFFF for x in buf]
        t0 = v[40]
        rows.append({"tick_us_events": round(e0.elapsed_time(e1) * 1e3, 1),
                     "decode": [round((v[40] - t0) / 1e3, 1), round((v[41] - t0) / 1e3, 1)],
                     "letterbox": [round((v[42] - t0) / 1e3, 1), round((v[43] - t0) / 1e3, 1)],
                     "nms(+tracker)": [round((v[44] - t0) / 1e3, 1), round((v[45] - t0) / 1e3, 1)],
                     "tracker": [round((v[46] - t0) / 1e3, 1), round((v[47] - t0) / 1e3, 1)]})
    out[f"schedule {schedule}"] = rows
    print(f"schedule {schedule}:")
    for r in rows:
        print("  ", json.dumps(r))
    h.close()

"""Phase stamps of the fused k_post_track (timing build) on a few-stream tick: STREAMS=4 python tools/phase_timing_fused.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["B200VA_LIB"] = os.path.join(ROOT, "realtime_video_analytics_32streams_b200", "lib", "libb200va_timing.so")
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench as B
from realtime_video_analytics_32streams_b200 import _native
S = int(os.environ.get("STREAMS", 4))
dev = torch.device("cuda", 0)
frames = torch.randint(0, 256, (S, B.H, B.W, 3), dtype=torch.uint8, device=dev)
heads_np = np.stack([B.make_heads(s, 3) for s in range(S)], axis=1)
heads = [torch.from_numpy(heads_np[k]).to(dev) for k in range(3)]
metas = (_native.Letterbox * S)(*[_native.letterbox_meta(B.H, B.W, *B.IN_HW) for _ in range(S)])
h = _native.Handle(device=0, max_batch=S, max_anchors=B.A, max_candidates=2048, max_dets=512, max_streams=S, max_tracks=1024)
h.lib.b200va_debug_read.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_int]
net = torch.empty((S, 3, *B.IN_HW), dtype=torch.float32, device=dev)
dets, tracks = h.alloc_dets(S), h.alloc_tracks(S)
fb = _native.FrameBatch(list(frames.unbind(0)))
plans = [h.plan_tick(frames=fb, net_out=net, dst_hw=B.IN_HW, head=heads[k], metas=metas, conf_thr=B.CONF, iou_thr=B.IOU,
                     filter_conf=B.CONF, dets=dets, slots=list(range(S)), tracker_cfg=(30, 1, 0.5), tracks=tracks, schedule=int(os.environ.get('SCHEDULE', 1)))
         for k in range(3)]
for rep in range(3):
    for k in range(6):
        h.tick(plans[k % 3])
    torch.cuda.synchronize()
    buf = (C.c_int64 * 64)()
    h.lib.b200va_debug_read(h._h, buf, 64)
    v = list(buf)
    seq = [16, 50, 17, 18, 19, 51, 52, 53, 20, 54, 55, 56, 21]
    names = ["loads+bar", "init+bar", "sort+bar", "gather+bar", "pairs", "bar", "resolve", "bar", "filter", "tid0", "bar", "emit"]
    print("nms:", ", ".join(f"{n} {v[b] - v[a]}" for n, a, b in zip(names, seq[:-1], seq[1:])), "| total", v[21] - v[16])
    print("nms end -> tracker start", v[0] - v[21])
    print("tracker: table staged %d, dets staged %d, phase A(+A2) %d, phase B+C %d, prune %d, fence+sync %d, ticket %d | total %d" %
          (v[1] - v[0], v[2] - v[1], v[7] - v[2], v[3] - v[7], v[4] - v[3], v[5] - v[4], v[6] - v[5], v[6] - v[0]))
    print("tracker phase A: zero+bar %d, pass1 %d, bar %d, pass2 %d, bar %d, A2 %d" % (v[58] - v[2], v[59] - v[58], v[60] - v[59], v[61] - v[60], v[62] - v[61], v[7] - v[62]))
    print("dets", dets["count"].cpu().tolist(), "tracks", tracks["count"].cpu().tolist())

"""Phase stamps of the fused k_post_track (timing build) on a few-stream tick: STREAMS=4 python tools/phase_timing_fused.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["B200VA_LIB"] = os.path.join(ROOT, "realtime_video_analytics_32streams_b200", "lib", "libb200va.so")
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
for k in range(12):
    h.tick(plans[k % 3])
torch.cuda.synchronize()
print('ok', dets['count'].cpu().tolist())

"""Phase timing of the latency-bound kernels (needs lib/libb200va_timing.so: make -C csrc TIMING=1)."""
import ctypes as C, os, sys
os.environ["B200VA_LIB"] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                        "realtime_video_analytics_32streams_b200", "lib", os.environ.get("TIMING_LIB", "libb200va_timing.so"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as B
from realtime_video_analytics_32streams_b200 import _native
h = _native.Handle(device=0, max_batch=32, max_anchors=B.A, max_candidates=4096, max_dets=1024, max_streams=64, max_tracks=2048)
if os.environ.get("DENSE"):
    from realtime_video_analytics_32streams_b200 import synth
    sc = [synth.DenseScene(9000 + s, n_objects=300, dup=6, n_obj_classes=10) for s in range(32)]
    heads = torch.from_numpy(np.stack([np.stack([c.head(t) for c in sc]) for t in range(2)])).cuda()
else:
    heads = torch.from_numpy(np.stack([B.make_heads(s, 2) for s in range(32)], axis=1)).cuda()
metas = (_native.Letterbox * 32)(*[_native.letterbox_meta(B.H, B.W, 640, 640) for _ in range(32)])
dets, tracks = h.alloc_dets(32), h.alloc_tracks(32)
for k in range(6):
    h.postprocess(heads[k % 2], metas, B.CONF, B.IOU, filter_conf=B.CONF, out=dets)
    h.tracker_update(list(range(32)), dets, 30, 1, 0.5, out=tracks)
buf = (C.c_int64 * 64)()
h.lib.b200va_debug_read.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_int]
h.lib.b200va_debug_read(h._h, buf, 64)
v = list(buf)
print("tracker phases (SM cycles, block 0): start->staged-issue %d, ->dets staged %d, ->match loop done %d, ->prune done %d, ->fence+sync %d, ->ticket %d" %
      tuple(v[i + 1] - v[i] for i in range(6)))
print("tracker: phase A %d, phase B %d; first iterations of B: %s" % (v[7] - v[2], v[3] - v[7], [v[9 + i] - v[8 + i] for i in range(7)]))
print("tracker phase A split (last chunk): clear %d, scan %d, barrier %d, iou %d, barrier %d, classify %d; n_plist?" % (v[58] - v[2], v[59] - v[58], v[60] - v[59], v[61] - v[60], v[62] - v[61], v[7] - v[62]))
print("nms phases: count+keys %d, sort %d, gather %d, chunks %d, filter+emit %d" % tuple(v[16 + i + 1] - v[16 + i] for i in range(5)))
print("nms chunk loop split: (q) grid query (warp 0) %d, (a) pair matrix %d, (b) resolve %d, (c) tail / insert %d" % (v[27], v[24], v[25], v[26]))
print("grid query split (accumulated over launches, thread 0): setup %d, cells %d, overflow %d, pass2 %d" % (v[32], v[33], v[34], v[35]))
print("grid: max entries per cell %d, overflow list %d, total entries %d" % (v[28], v[29], v[30]))
print("dense resolve (SM cycles, frame 0): loads+lists %d, Jacobi rounds %d, compaction %d, rank+emit %d" % (v[37] - v[36], v[38] - v[37], v[39] - v[38], v[48] - v[39]))
print("dets per frame", dets["count"].cpu().tolist()[:8], "tracks", tracks["count"].cpu().tolist()[:8])

# hypothesis check: does a low-occupancy grid (32 CTAs) run slower per instruction than when the rest of the chip is busy?
side = torch.cuda.Stream()
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
for busy in (False, True):
    torch.cuda.synchronize()
    if busy:
        with torch.cuda.stream(side):
            for _ in range(40):
                a @ a
    for k in range(4):
        h.postprocess(heads[k % 2], metas, B.CONF, B.IOU, filter_conf=B.CONF, out=dets)
        h.tracker_update(list(range(32)), dets, 30, 1, 0.5, out=tracks)
    h.lib.b200va_debug_read(h._h, buf, 64)
    v = list(buf)
    print("busy" if busy else "idle", "tracker A %d B %d  per-iter %s | nms sort %d chunks %d" %
          (v[7] - v[2], v[3] - v[7], [v[9 + i] - v[8 + i] for i in range(3)], v[18] - v[17], v[20] - v[19]))

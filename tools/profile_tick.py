"""Host-side time breakdown of HotPathEngine.tick on the bench workload (developer tool)."""
import cProfile, pstats, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as B
from realtime_video_analytics_32streams_b200 import DetectorConfig, HotPathEngine, StreamConfig, TrackerConfig, _native

dev = torch.device("cuda", 0)
h = _native.Handle(device=0, max_batch=32, max_anchors=B.A, max_candidates=2048, max_dets=512, max_streams=64, max_tracks=1024)
heads_np = np.stack([B.make_heads(s, B.N_SETS) for s in range(B.STREAMS)], axis=1)
host_heads = [torch.from_numpy(heads_np[k]).pin_memory() for k in range(B.N_SETS)]
host_frames = [torch.randint(0, 256, (B.H, B.W, 3), dtype=torch.uint8).pin_memory() for _ in range(B.STREAMS)]
tick = [0]
def infer(t): return host_heads[tick[0] % B.N_SETS].to(dev, non_blocking=True)
streams = [StreamConfig(name=f"s{s}") for s in range(B.STREAMS)]
for objs in (True, False):
    eng = HotPathEngine(streams, DetectorConfig(confidence_threshold=B.CONF, iou_threshold=B.IOU), TrackerConfig(**B.TRK), infer=infer, handle=h, input_hw=B.IN_HW, build_objects=objs)
    for k in range(5):
        tick[0] = k; eng.tick(host_frames)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(50):
        tick[0] = k; eng.tick(host_frames)
    torch.cuda.synchronize(); print("objects", objs, "ms/tick", (time.perf_counter() - t0) / 50 * 1e3)
# raw H2D rates
big = torch.empty(156_672_000, dtype=torch.uint8).pin_memory(); dbig = torch.empty_like(big, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): dbig.copy_(big, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
print("contiguous H2D 156.7MB: %.3f ms  %.1f GB/s" % (dt * 1e3, 156.672 / dt / 1e3))
st = eng.stager
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): st.upload(host_frames, sparse_for=B.IN_HW)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
print("sparse upload 32x1080p (66.4MB): %.3f ms  %.1f GB/s" % (dt * 1e3, 66.355 / dt / 1e3))
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): st.upload(host_frames)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
print("full upload 32x1080p (199MB): %.3f ms  %.1f GB/s" % (dt * 1e3, 199.07 / dt / 1e3))
eng = HotPathEngine(streams, DetectorConfig(confidence_threshold=B.CONF, iou_threshold=B.IOU), TrackerConfig(**B.TRK), infer=infer, handle=h, input_hw=B.IN_HW, build_objects=True)
for k in range(3): eng.tick(host_frames)
pr = cProfile.Profile(); pr.enable()
for k in range(30):
    tick[0] = k; eng.tick(host_frames)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)

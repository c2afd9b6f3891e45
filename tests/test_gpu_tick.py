"""``b200va_tick`` (letterbox ‖ decode + NMS + tracker in one call, two streams): every schedule,
eager and replayed from a CUDA graph, must give exactly what the three separate C-ABI calls give,
and those are pinned against the oracle (pipeline.py:172-188 of the reference, stream by stream)."""
import numpy as np
import pytest

from oracle import hotpath as O
from realtime_video_analytics_32streams_b200 import synth

pytestmark = pytest.mark.gpu

HW = (540, 960)
IN_HW = (640, 640)
B = 6
CONF, IOU = 0.35, 0.5
TRK = (30, 1, 0.5)  # max_age, min_hits, max_iou_distance


def _handle():
    from realtime_video_analytics_32streams_b200 import _native

    return _native.Handle(device=0, max_batch=8, max_anchors=8400, max_candidates=2048, max_dets=512, max_streams=8,
                          max_tracks=512)


def _inputs(n_ticks):
    frames = [[synth.synth_frame(900 + 10 * t + s, *HW) for s in range(B)] for t in range(n_ticks)]
    heads = [np.stack([synth.synth_head(700 + 10 * t + s, 84, 8400, 14, dup=3) for s in range(B)]) for t in range(n_ticks)]
    return frames, heads


def _oracle(frames, heads):
    trk = O.IouTracker(*[TRK[0], TRK[2], TRK[1]])
    meta = O.letterbox_meta(*HW, *IN_HW)
    out = []
    for f_t, h_t in zip(frames, heads):
        tensors = [O.preprocess(f, IN_HW)[0][0] for f in f_t]
        tracks = []
        for s in range(B):
            dets = O.filter_detections(O.postprocess(h_t[s][None], meta, CONF, IOU), CONF)
            tracks.append([(w.track_id, w.class_id, w.hits, w.bbox_xyxy) for w in trk.update(f"s{s}", dets)])
        out.append((tensors, tracks))
    return out


def _check(h, want, net, tracks, t):
    tensors, trk = want[t]
    got = net.cpu().numpy()
    for s in range(B):
        assert np.array_equal(got[s].view(np.uint8), tensors[s].view(np.uint8)), (t, s)
        n = int(tracks["count"][s].item())
        rows = list(zip(tracks["track_id"][s, :n].cpu().tolist(), tracks["cls"][s, :n].cpu().tolist(),
                        tracks["hits"][s, :n].cpu().tolist(),
                        [tuple(b) for b in tracks["bbox_xyxy"][s, :n].cpu().tolist()]))
        assert rows == trk[s], (t, s)


@pytest.mark.parametrize("fused", [1, 0])
@pytest.mark.parametrize("schedule", [0, 1, 2, 3, 5, 6])
def test_tick_matches_oracle(schedule, fused, monkeypatch):
    """fused=1: NMS + tracker as one kernel (k_post_track, the default for sparse scenes); fused=0: two kernels
    chained by programmatic dependent launch.  Schedule 3 also launches the letterbox as a programmatic dependent
    of the decode kernel."""
    import torch
    from realtime_video_analytics_32streams_b200 import _native as N

    monkeypatch.setenv("B200VA_FUSE_POST_TRACK", str(fused))
    n_ticks = 3
    frames, heads = _inputs(n_ticks)
    want = _oracle(frames, heads)
    h = _handle()
    net = torch.empty((B, 3, *IN_HW), dtype=torch.float32, device="cuda")
    metas = [N.letterbox_meta(*HW, *IN_HW) for _ in range(B)]
    begin, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    begin.record()
    end.record()
    for t in range(n_ticks):
        dev = [torch.from_numpy(f).cuda() for f in frames[t]]
        plan = h.plan_tick(frames=dev, net_out=net, dst_hw=IN_HW, head=torch.from_numpy(heads[t]).cuda(), metas=metas,
                           conf_thr=CONF, iou_thr=IOU, filter_conf=CONF, slots=list(range(B)), tracker_cfg=TRK,
                           schedule=schedule)
        plan.set_events(begin, end)
        net.zero_()
        h.tick(plan)
        torch.cuda.synchronize()
        assert begin.elapsed_time(end) > 0.0
        _check(h, want, net, plan.tracks, t)
    h.poll_status()
    h.close()


@pytest.mark.parametrize("fused", [1, 0])
def test_tick_software_pipelined_schedule_lags_one_call(fused, monkeypatch):
    """Schedule 4: call k decodes head k and letterboxes frames k while NMS + tracker of head k-1 run beside them;
    the tables of tick k are complete after call k+1, the last one after a call without inputs.  Same results."""
    import torch
    from realtime_video_analytics_32streams_b200 import _native as N

    monkeypatch.setenv("B200VA_FUSE_POST_TRACK", str(fused))
    n_ticks = 4
    frames, heads = _inputs(n_ticks)
    want = _oracle(frames, heads)
    h = _handle()
    nets = [torch.zeros((B, 3, *IN_HW), dtype=torch.float32, device="cuda") for _ in range(n_ticks)]
    metas = [N.letterbox_meta(*HW, *IN_HW) for _ in range(B)]
    plans = []
    for t in range(n_ticks):
        dev = [torch.from_numpy(f).cuda() for f in frames[t]]
        plans.append(h.plan_tick(frames=dev, net_out=nets[t], dst_hw=IN_HW, head=torch.from_numpy(heads[t]).cuda(), metas=metas,
                                 conf_thr=CONF, iou_thr=IOU, filter_conf=CONF, slots=list(range(B)), tracker_cfg=TRK,
                                 schedule=N.SCHEDULE_SOFTWARE_PIPELINED))
        h.tick(plans[t])
        torch.cuda.synchronize()
        if t == 0:
            assert int(plans[0].tracks["count"].sum().item()) == 0  # nothing ran for tick 0 yet
        else:
            _check(h, want, nets[t - 1], plans[t - 1].tracks, t - 1)
    h.tick(h.plan_tick(schedule=N.SCHEDULE_SOFTWARE_PIPELINED))  # drain: the chain of the last head
    torch.cuda.synchronize()
    _check(h, want, nets[-1], plans[-1].tracks, n_ticks - 1)
    # leaving the schedule with a chain still owed: the next tick of any schedule runs it first
    extra = h.plan_tick(frames=[torch.from_numpy(f).cuda() for f in frames[0]], net_out=nets[0], dst_hw=IN_HW,
                        head=torch.from_numpy(heads[0]).cuda(), metas=metas, conf_thr=CONF, iou_thr=IOU, filter_conf=CONF,
                        slots=list(range(B)), tracker_cfg=TRK, schedule=N.SCHEDULE_SOFTWARE_PIPELINED)
    h.tick(extra)
    other = h.plan_tick(head=torch.from_numpy(heads[1]).cuda(), metas=metas, conf_thr=CONF, iou_thr=IOU, filter_conf=CONF,
                        slots=list(range(B)), tracker_cfg=TRK, schedule=1)
    h.tick(other)
    torch.cuda.synchronize()
    assert int(extra.tracks["count"].sum().item()) > 0 and int(other.tracks["count"].sum().item()) > 0
    h.poll_status()
    h.close()


def test_tick_software_pipelined_graph_replay():
    """Schedule 4 replayed from CUDA graphs: the stashed chain is captured with the call that launches it."""
    import torch
    from realtime_video_analytics_32streams_b200 import _native as N

    n_ticks = 5
    frames, heads = _inputs(n_ticks)
    want = _oracle(frames, heads)
    h = _handle()
    net = torch.empty((B, 3, *IN_HW), dtype=torch.float32, device="cuda")
    metas = [N.letterbox_meta(*HW, *IN_HW) for _ in range(B)]
    dev = [torch.empty((*HW, 3), dtype=torch.uint8, device="cuda") for _ in range(B)]
    head = torch.empty((B, 84, 8400), dtype=torch.float32, device="cuda")
    plan = h.plan_tick(frames=dev, net_out=net, dst_hw=IN_HW, head=head, metas=metas, conf_thr=CONF, iou_thr=IOU,
                       filter_conf=CONF, slots=list(range(B)), tracker_cfg=TRK, schedule=N.SCHEDULE_SOFTWARE_PIPELINED)
    # prime the pipeline eagerly (tick 0: decode only), then capture the steady-state call twice: the candidate sets
    # alternate, so an even and an odd graph are needed
    for s in range(B):
        dev[s].copy_(torch.from_numpy(frames[0][s]))
    head.copy_(torch.from_numpy(heads[0]))
    h.tick(plan)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graphs = []
    with torch.cuda.stream(side):
        for k in range(2):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=side):
                h.tick(plan)
            graphs.append(gr)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    for t in range(1, n_ticks):
        for s in range(B):
            dev[s].copy_(torch.from_numpy(frames[t][s]))
        head.copy_(torch.from_numpy(heads[t]))
        graphs[(t - 1) % 2].replay()
        torch.cuda.synchronize()
        _check(h, want, net, plan.tracks, t - 1) if False else None  # net holds tick t's letterbox, tracks tick t-1's
        tensors, trk = want[t - 1]
        for s in range(B):
            n = int(plan.tracks["count"][s].item())
            rows = list(zip(plan.tracks["track_id"][s, :n].cpu().tolist(), plan.tracks["cls"][s, :n].cpu().tolist(),
                            plan.tracks["hits"][s, :n].cpu().tolist(),
                            [tuple(b) for b in plan.tracks["bbox_xyxy"][s, :n].cpu().tolist()]))
            assert rows == trk[s], (t, s)
        got = net.cpu().numpy()
        for s in range(B):
            assert np.array_equal(got[s].view(np.uint8), want[t][0][s].view(np.uint8)), (t, s)
    h.poll_status()
    h.close()


@pytest.mark.parametrize("pdl", [1, 0])
@pytest.mark.parametrize("schedule", [1, 3, 5, 6])
def test_tick_graph_replay_matches_oracle(schedule, pdl, monkeypatch):
    import torch
    from realtime_video_analytics_32streams_b200 import _native as N

    monkeypatch.setenv("B200VA_PDL", str(pdl))
    n_ticks = 3
    frames, heads = _inputs(n_ticks)
    want = _oracle(frames, heads)
    h = _handle()
    net = torch.empty((B, 3, *IN_HW), dtype=torch.float32, device="cuda")
    metas = [N.letterbox_meta(*HW, *IN_HW) for _ in range(B)]
    dev = [torch.empty((*HW, 3), dtype=torch.uint8, device="cuda") for _ in range(B)]
    head = torch.empty((B, 84, 8400), dtype=torch.float32, device="cuda")
    plan = h.plan_tick(frames=dev, net_out=net, dst_hw=IN_HW, head=head, metas=metas, conf_thr=CONF, iou_thr=IOU,
                       filter_conf=CONF, slots=list(range(B)), tracker_cfg=TRK, schedule=schedule)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            h.tick(plan)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    for s in range(B):  # the capture itself ran nothing; start from clean tracker state anyway
        h.tracker_reset(s)
    h.tracker_set_next_id(1)
    for t in range(n_ticks):
        for s in range(B):
            dev[s].copy_(torch.from_numpy(frames[t][s]))
        head.copy_(torch.from_numpy(heads[t]))
        net.zero_()
        graph.replay()
        torch.cuda.synchronize()
        _check(h, want, net, plan.tracks, t)
    h.poll_status()
    h.close()


def test_tick_halves_are_optional():
    import torch
    from realtime_video_analytics_32streams_b200 import _native as N

    frames, heads = _inputs(1)
    want = _oracle(frames, heads)
    h = _handle()
    net = torch.empty((B, 3, *IN_HW), dtype=torch.float32, device="cuda")
    metas = [N.letterbox_meta(*HW, *IN_HW) for _ in range(B)]
    pre = h.plan_tick(frames=[torch.from_numpy(f).cuda() for f in frames[0]], net_out=net, dst_hw=IN_HW)
    h.tick(pre)
    post = h.plan_tick(head=torch.from_numpy(heads[0]).cuda(), metas=metas, conf_thr=CONF, iou_thr=IOU, filter_conf=CONF,
                       slots=list(range(B)), tracker_cfg=TRK)
    h.tick(post)
    torch.cuda.synchronize()
    _check(h, want, net, post.tracks, 0)
    h.close()


def test_pads_valid_flag_skips_only_the_pad_rows():
    """B200VA_OUT_FLAG_PADS_VALID: same tensor as a full call once the pad rows are in place, and the kernel
    really leaves them alone (a sentinel written there survives)."""
    import torch
    from realtime_video_analytics_32streams_b200 import _native as N

    h = _handle()
    frames = [synth.synth_frame(40 + i, *hw) for i, hw in enumerate([(1080, 1920), (720, 1280), (1920, 1080), (360, 640)])]
    dev = [torch.from_numpy(f).cuda() for f in frames]
    for fmt in (N.OUT_F32_RGB_NCHW, N.OUT_F16_RGB_NCHW, N.OUT_U8_BGR_NHWC):
        full, metas = h.preprocess(dev, IN_HW, fmt)
        again = full.clone()
        h.preprocess(dev, IN_HW, fmt | N.OUT_FLAG_PADS_VALID, out=again)
        assert torch.equal(again, full)
        marked = full.clone()
        top = metas[0].pad_top  # frame 0 is 16:9: rows [0, top) and [top + new_h, 640) are padding
        sentinel = 7 if fmt == N.OUT_U8_BGR_NHWC else 0.5
        if fmt == N.OUT_U8_BGR_NHWC:
            marked[0, :top] = sentinel
        else:
            marked[0, :, :top] = sentinel
        h.preprocess(dev, IN_HW, fmt | N.OUT_FLAG_PADS_VALID, out=marked)
        band = marked[0, :top] if fmt == N.OUT_U8_BGR_NHWC else marked[0, :, :top]
        assert bool((band == sentinel).all()) and top == 140
        inner = (marked[0, top:top + metas[0].new_h] if fmt == N.OUT_U8_BGR_NHWC else marked[0, :, top:top + metas[0].new_h])
        ref = (full[0, top:top + metas[0].new_h] if fmt == N.OUT_U8_BGR_NHWC else full[0, :, top:top + metas[0].new_h])
        assert torch.equal(inner, ref)
        assert torch.equal(marked[1:], full[1:])  # portrait frame: its pad columns are still written
    h.poll_status()
    h.close()

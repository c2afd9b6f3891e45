"""GPU property tests at BASELINE.json's full sizes (32 streams, 1080p / 4K, [32,84,8400] heads),
where running the NumPy oracle on every frame would take minutes: size-independent properties plus
spot checks of single frames against the committed golden digests.  All calls go through the C ABI."""
import numpy as np
import pytest

import golden_util as G
from oracle import hotpath as O
from realtime_video_analytics_32streams_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def H():
    import torch
    from realtime_video_analytics_32streams_b200 import _native

    assert torch.cuda.is_available()
    h = _native.Handle(device=0, max_batch=64, max_anchors=8400, max_candidates=4096, max_dets=1024, max_streams=64,
                       max_tracks=2048)
    yield h
    h.poll_status()
    h.close()


def test_preprocess_32x1080p_batch_equals_single_frame_calls_and_golden_digest(H):
    """Batch independence: frame b of a 32-frame launch == the same frame launched alone; frame 0 is
    the golden 1080p frame, so the whole batch is anchored to the reference's own output."""
    import torch
    from realtime_video_analytics_32streams_b200 import _native as N

    d = G.meta()["preprocess_digests"]["1080p_f32"]
    g = torch.Generator(device="cuda")
    g.manual_seed(11)
    frames = torch.randint(0, 256, (32, 1080, 1920, 3), dtype=torch.uint8, device="cuda", generator=g)
    frames[0] = torch.from_numpy(synth.synth_frame(d["seed"], 1080, 1920)).cuda()
    batch, _ = H.preprocess(list(frames.unbind(0)), (640, 640), N.OUT_F32_RGB_NCHW)
    assert G.sha(batch[0:1].cpu().numpy()) == d["sha256"]
    for b in (1, 7, 31):
        one, _ = H.preprocess([frames[b]], (640, 640), N.OUT_F32_RGB_NCHW)
        assert torch.equal(one[0], batch[b])
    # 1080p -> 640x360 is a pure subsample of rows/columns 3k+1 (SURVEY.md a1'): check the whole batch
    sub = frames[:, 1::3, 1::3, :].flip(-1).permute(0, 3, 1, 2).float() * torch.tensor(1.0 / 255.0, dtype=torch.float32)
    assert torch.equal(batch[:, :, 140:500, :], sub.cuda())
    pad = torch.tensor(114.0, dtype=torch.float32) * torch.tensor(1.0 / 255.0, dtype=torch.float32)
    assert bool((batch[:, :, :140] == pad.item()).all()) and bool((batch[:, :, 500:] == pad.item()).all())


def test_preprocess_32x4k_is_2x2_box_average(H):
    """4K -> 640x360: every output pixel is (2x2 sum at rows 6d+2..3, cols 6d+2..3, + 2) >> 2."""
    import torch
    from realtime_video_analytics_32streams_b200 import _native as N

    g = torch.Generator(device="cuda")
    g.manual_seed(12)
    for chunk in range(2):  # 2 x 16 frames of 4K: 400 MB each
        frames = torch.randint(0, 256, (16, 2160, 3840, 3), dtype=torch.uint8, device="cuda", generator=g)
        out, _ = H.preprocess(list(frames.unbind(0)), (640, 640), N.OUT_U8_BGR_NHWC)
        f = frames.to(torch.int32)
        box = (f[:, 2::6, 2::6] + f[:, 2::6, 3::6] + f[:, 3::6, 2::6] + f[:, 3::6, 3::6] + 2) >> 2
        assert torch.equal(out[:, 140:500].to(torch.int32), box)
        assert bool((out[:, :140] == 114).all()) and bool((out[:, 500:] == 114).all())
    d = G.meta()["preprocess_digests"]["4k_f32"]
    one, _ = H.preprocess([torch.from_numpy(synth.synth_frame(d["seed"], 2160, 3840)).cuda()], (640, 640), N.OUT_F32_RGB_NCHW)
    assert G.sha(one.cpu().numpy()) == d["sha256"]


def test_postprocess_full_batch_properties(H):
    """[32, 84, 8400] dense heads: kept scores are sorted, kept boxes are mutually non-suppressing,
    every dropped candidate is suppressed by an earlier kept box (greedy NMS invariant), NMS is
    idempotent on its own output, and frame 0 matches the golden dense digest."""
    import torch
    from realtime_video_analytics_32streams_b200 import _native as N

    scenes = [synth.DenseScene(5 + s) for s in range(32)]
    heads = np.stack([sc.head(0) for sc in scenes])
    lb = [N.letterbox_meta(1080, 1920, 640, 640)] * 32
    out = H.postprocess(torch.from_numpy(heads).cuda(), lb, 0.35, 0.5)
    cnt = out["count"].cpu().numpy()
    d = G.meta()["postprocess_digests"]["dense_seed5_t0"]
    n0 = int(cnt[0])
    assert n0 == d["n"]
    assert G.sha(out["bbox_xyxy"][0, :n0].cpu().numpy().astype(np.float64)) == d["box"]
    meta = O.letterbox_meta(1080, 1920, 640, 640)
    thr = np.float32(0.5)
    for b in (3, 17, 31):
        n = int(cnt[b])
        conf = out["conf"][b, :n].cpu().numpy()
        box = out["bbox_xyxy"][b, :n].cpu().numpy()
        assert n > 250 and np.all(conf[:-1] > conf[1:])
        iou = O.pairwise_iou_f32(box)
        np.fill_diagonal(iou, 0)
        assert np.all(iou <= thr)  # survivors never suppress each other
        xywh, cconf, ccls, _ = O.decode_candidates(heads[b][None], 0.35, None)
        cbox = O.scale_boxes(O.xywh2xyxy(xywh), meta)
        kept = {(float(c), tuple(bx.tolist())) for c, bx in zip(conf, box)}
        dropped = [i for i in range(len(cconf)) if (float(cconf[i]), tuple(cbox[i].tolist())) not in kept]
        assert len(dropped) + n == len(cconf)
        for i in dropped[:200]:
            better = box[conf > cconf[i]]
            assert better.size and np.any(~(O.pairwise_iou_f32(np.vstack([cbox[i:i + 1], better]))[0, 1:] <= thr))
    # idempotence: feed the kept boxes of frame 3 back as a C == 5 head; nothing more is removed
    n = int(cnt[3])
    box = out["bbox_xyxy"][3, :n].cpu().numpy()
    conf = out["conf"][3, :n].cpu().numpy()
    head = np.zeros((1, 5, n), np.float32)
    head[0, 0] = (box[:, 0] + box[:, 2]) / 2
    head[0, 1] = (box[:, 1] + box[:, 3]) / 2
    head[0, 2] = box[:, 2] - box[:, 0]
    head[0, 3] = box[:, 3] - box[:, 1]
    head[0, 4] = conf
    ident = N.Letterbox(1080, 1920, 1080, 1920, 0, 0, 1.0)
    again = H.postprocess(torch.from_numpy(head).cuda(), [ident], 0.35, 0.5)
    assert int(again["count"].cpu()[0]) == n


def test_tracker_batch_composition_does_not_change_per_stream_results(H):
    """32 streams in one launch == the same streams in launches of 8: per-stream tables (boxes, hits,
    ages, order) are identical; only the globally numbered ids differ, by a per-tick permutation."""
    import torch
    from realtime_video_analytics_32streams_b200 import _native as N

    scenes = [synth.DenseScene(700 + s, n_objects=40, dup=3, n_obj_classes=5) for s in range(32)]
    lb = [N.letterbox_meta(1080, 1920, 640, 640)] * 32
    results = {}
    for mode, groups in (("all", [list(range(32))]), ("by8", [list(range(k, k + 8)) for k in range(0, 32, 8)])):
        H.tracker_set_next_id(1)
        for s in range(32):
            H.tracker_reset(s)
        seq = []
        for t in range(6):
            heads = torch.from_numpy(np.stack([sc.head(t) for sc in scenes])).cuda()
            dets = H.postprocess(heads, lb, 0.35, 0.5, filter_conf=0.35)
            per_stream = {}
            for grp in groups:
                sub = {k: v[grp[0]:grp[-1] + 1] for k, v in dets.items() if not k.startswith("_")}
                out = H.tracker_update(grp, sub, 30, 1, 0.5)
                cnt = out["count"].cpu().numpy()
                for j, s in enumerate(grp):
                    n = int(cnt[j])
                    per_stream[s] = {k: out[k][j, :n].cpu().numpy() for k in ("track_id", "cls", "conf", "bbox_xyxy", "age", "hits")}
            seq.append(per_stream)
        results[mode] = seq
    for t in range(6):
        a, b = results["all"][t], results["by8"][t]
        ids_a, ids_b = [], []
        for s in range(32):
            for k in ("cls", "conf", "bbox_xyxy", "age", "hits"):
                assert np.array_equal(a[s][k], b[s][k]), (t, s, k)
            ids_a.append(a[s]["track_id"])
            ids_b.append(b[s]["track_id"])
        ia, ib = np.concatenate(ids_a), np.concatenate(ids_b)
        assert len(set(ia.tolist())) == len(ia) and sorted(ia.tolist()) == sorted(ib.tolist())
        assert np.array_equal(ia, ib)  # same canonical stream order in both modes -> identical ids


def test_motion_32x1080p_static_scene_counts_zero_and_state_is_idempotent(H):
    import torch

    g = torch.Generator(device="cuda")
    g.manual_seed(13)
    frames = torch.randint(0, 256, (32, 1080, 1920, 3), dtype=torch.uint8, device="cuda", generator=g)
    fl = list(frames.unbind(0))
    a = [torch.empty((1080, 1920), dtype=torch.uint8, device="cuda") for _ in range(32)]
    b = [torch.empty((1080, 1920), dtype=torch.uint8, device="cuda") for _ in range(32)]
    assert (H.motion(fl, [None] * 32, a).cpu().numpy() == -1).all()
    assert (H.motion(fl, a, b).cpu().numpy() == 0).all()  # same frames: nothing changed
    for x, y in zip(a, b):
        assert torch.equal(x, y)  # the blurred gray is a pure function of the frame
    mf = O.MotionFilter()
    mf.should_process(frames[5].cpu().numpy())
    assert np.array_equal(a[5].cpu().numpy(), mf.previous_gray)
    # inverting a block changes exactly the pixels the oracle says
    frames[9, 200:500, 300:900] = 255 - frames[9, 200:500, 300:900]
    cnt = H.motion(fl, b, a).cpu().numpy()
    assert (np.delete(cnt, 9) == 0).all()
    omf = O.MotionFilter()
    g9 = frames[9].cpu().numpy().copy()
    g9[200:500, 300:900] = 255 - g9[200:500, 300:900]
    omf.should_process(g9)
    omf.should_process(frames[9].cpu().numpy())
    assert int(cnt[9]) == omf.last_count


def test_tick_32x1080p_graph_replay_equals_separate_calls():
    """BASELINE config 3 shape through the bench's own path: the prepared b200va_tick replayed from a CUDA graph
    (schedule 1, with and without B200VA_OUT_FLAG_PADS_VALID) against the three separate C-ABI calls on a second
    handle -- network input, detections and track tables, three ticks, sparse and dense heads."""
    import torch
    from realtime_video_analytics_32streams_b200 import _native as N

    B, hw, in_hw = 32, (1080, 1920), (640, 640)
    mk = lambda: N.Handle(device=0, max_batch=B, max_anchors=8400, max_candidates=2048, max_dets=512, max_streams=B,
                          max_tracks=1024)
    ha, hb = mk(), mk()
    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    frames = torch.randint(0, 256, (B, *hw, 3), dtype=torch.uint8, device="cuda", generator=g)
    fb = N.FrameBatch(list(frames.unbind(0)))
    head = torch.empty((B, 84, 8400), dtype=torch.float32, device="cuda")
    metas = (N.Letterbox * B)(*[N.letterbox_meta(*hw, *in_hw) for _ in range(B)])
    net_a = torch.zeros((B, 3, *in_hw), dtype=torch.float32, device="cuda")
    slots = list(range(B))
    for pads_flag in (0, N.OUT_FLAG_PADS_VALID):
        for s in slots:
            ha.tracker_reset(s)
            hb.tracker_reset(s)
        ha.tracker_set_next_id(1)
        hb.tracker_set_next_id(1)
        if pads_flag:
            ha.preprocess(fb, in_hw, N.OUT_F32_RGB_NCHW, out=net_a)  # the pad rows the flag relies on
        plan = ha.plan_tick(frames=fb, net_out=net_a, dst_hw=in_hw, fmt=N.OUT_F32_RGB_NCHW | pads_flag, head=head,
                            metas=metas, conf_thr=0.35, iou_thr=0.5, filter_conf=0.35, slots=slots,
                            tracker_cfg=(30, 1, 0.5), schedule=1)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                ha.tick(plan)
        torch.cuda.current_stream().wait_stream(side)
        for t, (n_obj, dup) in enumerate([(24, 3), (300, 6), (24, 3)]):
            frames.copy_(torch.randint(0, 256, frames.shape, dtype=torch.uint8, device="cuda", generator=g))
            scenes = [synth.DenseScene(8000 + s, n_objects=n_obj, dup=dup, n_obj_classes=10) for s in range(B)]
            head.copy_(torch.from_numpy(np.stack([sc.head(t) for sc in scenes])))
            graph.replay()
            net_b, _ = hb.preprocess(fb, in_hw, N.OUT_F32_RGB_NCHW)
            dets_b = hb.postprocess(head, metas, 0.35, 0.5, filter_conf=0.35)
            trk_b = hb.tracker_update(slots, dets_b, 30, 1, 0.5)
            torch.cuda.synchronize()
            assert torch.equal(net_a, net_b), (pads_flag, t)
            assert torch.equal(plan.dets["count"], dets_b["count"]), (pads_flag, t)
            live = torch.arange(plan.dets["conf"].shape[1], device="cuda")[None, :] < dets_b["count"][:, None]
            for k in ("cls", "conf", "bbox_xyxy"):  # rows past the count are stale scratch
                m = live if plan.dets[k].dim() == 2 else live[:, :, None].expand_as(plan.dets[k])
                assert torch.equal(plan.dets[k][m], dets_b[k][m]), (pads_flag, t, k)
            cnt = trk_b["count"]
            assert torch.equal(plan.tracks["count"], cnt)
            for k in ("track_id", "cls", "conf", "bbox_xyxy", "age", "hits"):
                for s in (0, 13, 31):
                    n = int(cnt[s])
                    assert torch.equal(plan.tracks[k][s, :n], trk_b[k][s, :n]), (pads_flag, t, k, s)
    ha.poll_status()
    hb.poll_status()
    ha.close()
    hb.close()

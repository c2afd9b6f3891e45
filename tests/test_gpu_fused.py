"""b200va_motion_preprocess: the motion gate and the letterbox in one pass over each frame must equal the two
separate entry points -- and through them the oracle (MotionFilter frame_filter.py:26-40, _preprocess
detector.py:198-264, apply_roi frame_filter.py:43-50) -- bit for bit."""
import numpy as np
import pytest

from oracle import hotpath as O
from realtime_video_analytics_32streams_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def H():
    import torch
    from realtime_video_analytics_32streams_b200 import _native

    assert torch.cuda.is_available()
    h = _native.Handle(device=0, max_batch=40, max_anchors=8400, max_candidates=1024, max_dets=256, max_streams=4,
                       max_tracks=64)
    yield h
    h.poll_status()
    h.close()


def cu(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


SHAPES = [(1080, 1920), (2160, 3840), (720, 1280), (1440, 2560), (360, 640), (1920, 1080), (1088, 1920), (368, 640),
          (1083, 1921),   # odd size: not tile-eligible, runs the separate kernels inside the call
          (300, 400),     # smaller than the network input: up-scaling geometry, not fusable
          (64, 4096), (2000, 16)]


@pytest.mark.parametrize("half", [False, True])
@pytest.mark.parametrize("with_masks", [False, True])
def test_fused_equals_oracle_two_frames(H, half, with_masks):
    import torch
    from realtime_video_analytics_32streams_b200 import _native as N

    fmt = N.OUT_F16_RGB_NCHW if half else N.OUT_F32_RGB_NCHW
    frames0 = [synth.synth_frame(500 + i, h, w) for i, (h, w) in enumerate(SHAPES)]
    frames1 = [np.clip(f.astype(np.int16) + np.random.default_rng(600 + i).integers(-40, 41, f.shape), 0, 255).astype(np.uint8)
               for i, f in enumerate(frames0)]
    polys = [synth.synth_polygons(700 + i, h, w) if (with_masks and i % 3 != 2) else None for i, (h, w) in enumerate(SHAPES)]
    masks = [H.roi_rasterize(p, h, w) if p is not None else None for p, (h, w) in zip(polys, SHAPES)]
    n = len(SHAPES)
    gray = [[torch.empty(hw, dtype=torch.uint8, device="cuda") for hw in SHAPES] for _ in range(2)]
    filters = [O.MotionFilter() for _ in range(n)]
    for t, frames in enumerate((frames0, frames1)):
        dev = [cu(f) for f in frames]
        prev = [None] * n if t == 0 else gray[(t + 1) % 2]
        changed, net, metas = H.motion_preprocess(dev, prev, gray[t % 2], (640, 640), fmt, masks)
        changed = changed.cpu().numpy()
        got = net.cpu().numpy()
        for i, f in enumerate(frames):
            masked = O.apply_roi(f, polys[i]) if polys[i] is not None else f
            filters[i].should_process(masked)
            ref, meta = O.preprocess(masked, (640, 640), half)
            assert np.array_equal(got[i].view(np.uint8), ref[0].view(np.uint8)), (t, SHAPES[i])
            assert metas[i].as_meta() == meta
            assert np.array_equal(gray[t % 2][i].cpu().numpy(), filters[i].previous_gray), (t, SHAPES[i])
            assert int(changed[i]) == (-1 if t == 0 else filters[i].last_count), (t, SHAPES[i])


def test_fused_equals_separate_calls_full_batch(H):
    """32 x 1080p and 32 x 4K with masks: identical to b200va_motion + b200va_preprocess; PADS_VALID keeps pad rows."""
    import torch
    from realtime_video_analytics_32streams_b200 import _native as N

    for (h, w), B in (((1080, 1920), 32), ((2160, 3840), 34)):
        g = torch.Generator(device="cuda")
        g.manual_seed(5)
        frames = list(torch.randint(0, 256, (B, h, w, 3), dtype=torch.uint8, device="cuda", generator=g).unbind(0))
        masks = [H.roi_rasterize(synth.synth_polygons(40 + s % 5, h, w), h, w) if s % 2 else None for s in range(B)]
        prev = [torch.randint(0, 256, (h, w), dtype=torch.uint8, device="cuda", generator=g) for _ in range(B)]
        nxt_a = [torch.empty((h, w), dtype=torch.uint8, device="cuda") for _ in range(B)]
        nxt_b = [torch.empty((h, w), dtype=torch.uint8, device="cuda") for _ in range(B)]
        want_changed = H.motion(frames, prev, nxt_a, masks).clone()
        want_net, _ = H.preprocess(frames, (640, 640), N.OUT_F32_RGB_NCHW, masks)
        changed, net, _ = H.motion_preprocess(frames, prev, nxt_b, (640, 640), N.OUT_F32_RGB_NCHW, masks)
        assert torch.equal(changed, want_changed)
        assert torch.equal(net, want_net)
        for a, b in zip(nxt_a, nxt_b):
            assert torch.equal(a, b)
        marked = want_net.clone()
        marked[:, :, :100] = 0.25  # inside the top pad band of a 16:9 frame
        H.motion_preprocess(frames, prev, nxt_b, (640, 640), N.OUT_F32_RGB_NCHW | N.OUT_FLAG_PADS_VALID, masks, out=marked)
        assert bool((marked[:, :, :100] == 0.25).all()) and torch.equal(marked[:, :, 140:500], want_net[:, :, 140:500])

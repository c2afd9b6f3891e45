"""The NumPy restatements of OpenCV's integer kernels (oracle/cv_restate.py) against the
installed cv2 (the third-party dependency the reference calls).  CPU-only."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from oracle import cv_restate as R
from oracle import hotpath as O


RESIZE_CASES = [((1080, 1920), (640, 360)), ((2160, 3840), (640, 360)), ((720, 1280), (640, 360)),
                ((1440, 2560), (640, 360)), ((1088, 1920), (640, 362)), ((480, 640), (640, 480)),
                ((360, 640), (640, 360)), ((100, 37), (64, 51)), ((37, 100), (200, 90)), ((333, 517), (129, 77)),
                ((5, 7), (3, 2)), ((2, 2), (7, 9)), ((1, 1), (4, 4)), ((1080, 1920), (960, 540)),
                ((1080, 1920), (1344, 756)), ((1920, 1080), (360, 640))]


@pytest.mark.parametrize("src_hw,dst_wh", RESIZE_CASES)
def test_resize_linear_matches_cv2(src_hw, dst_wh):
    rng = np.random.default_rng(hash((src_hw, dst_wh)) % (2 ** 32))
    src = rng.integers(0, 256, size=(*src_hw, 3), dtype=np.uint8)
    ref = cv2.resize(src, dst_wh, interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(R.resize_linear_u8(src, *dst_wh), ref)


@pytest.mark.parametrize("hw", [(1080, 1920), (37, 53), (5, 5), (3, 9), (2, 2), (1, 7), (4, 1)])
def test_gray_and_blur_match_cv2(hw):
    rng = np.random.default_rng(hw[0] * 7 + hw[1])
    src = rng.integers(0, 256, size=(*hw, 3), dtype=np.uint8)
    g = cv2.cvtColor(src, cv2.COLOR_BGR2GRAY)
    assert np.array_equal(R.bgr2gray(src), g)
    assert np.array_equal(R.gaussian_blur5(g), cv2.GaussianBlur(g, (5, 5), 0))


def test_threshold_is_strict():
    a = np.arange(256, dtype=np.uint8)[None]
    b = np.full_like(a, 100)
    _, th = cv2.threshold(cv2.absdiff(a, b), 25, 255, cv2.THRESH_BINARY)
    assert int(np.count_nonzero(th)) == R.motion_changed_count(a, b)


@pytest.mark.parametrize("lo,hi,hw,n", [(0.0, 1.0, (60, 80), 150), (-0.5, 1.5, (60, 80), 250),
                                        (0.0, 1.0, (270, 480), 30), (-0.5, 1.5, (270, 480), 40),
                                        (-3.0, 4.0, (40, 50), 100)])
def test_fill_poly_matches_cv2(lo, hi, hw, n):
    h, w = hw
    rng = np.random.default_rng(int(1000 * (hi - lo)) + h)
    for _ in range(n):
        k = int(rng.integers(3, 9))
        pts = np.stack([rng.integers(int(lo * w), int(hi * w), k), rng.integers(int(lo * h), int(hi * h), k)], 1)
        ref = np.zeros((h, w), np.uint8)
        cv2.fillPoly(ref, [pts.astype(np.int32)], 255)
        assert np.array_equal(R.fill_poly_mask(h, w, [pts.tolist()]), ref), pts.tolist()


def test_fill_poly_degenerate():
    for pts in ([(5, 5)], [(5, 5), (20, 9)], [(3, 3), (3, 3), (3, 3)], [(0, 10), (30, 10), (15, 10)]):
        ref = np.zeros((30, 40), np.uint8)
        cv2.fillPoly(ref, [np.array(pts, np.int32)], 255)
        assert np.array_equal(R.fill_poly_mask(30, 40, [pts]), ref), pts


def test_line8_matches_cv2():
    rng = np.random.default_rng(3)
    h, w = 50, 70
    for _ in range(1500):
        p = [int(v) for v in rng.integers(-40, 110, 4)]
        ref = np.zeros((h, w), np.uint8)
        cv2.line(ref, (p[0], p[1]), (p[2], p[3]), 255, 1, cv2.LINE_8)
        got = np.zeros((h, w), np.uint8)
        for x, y in R.line8_points(w, h, *p):
            got[y, x] = 255
        assert np.array_equal(got, ref), p


def test_numpy_and_cv2_back_ends_agree_end_to_end():
    rng = np.random.default_rng(9)
    frame = rng.integers(0, 256, size=(270, 480, 3), dtype=np.uint8)
    polys = [[(30, 20), (400, 40), (450, 250), (60, 200)]]
    a = O.apply_roi(frame, polys, "numpy")
    assert np.array_equal(a, O.apply_roi(frame, polys, "cv2"))
    assert np.array_equal(O.downsample(a, 0.6, "numpy"), O.downsample(a, 0.6, "cv2"))
    t0, m0 = O.preprocess(a, (160, 160), False, "numpy")
    t1, m1 = O.preprocess(a, (160, 160), False, "cv2")
    assert np.array_equal(t0, t1) and m0 == m1
    f0, f1 = O.MotionFilter(0.02, "numpy"), O.MotionFilter(0.02, "cv2")
    for k in range(3):
        fr = np.roll(frame, 7 * k, axis=1)
        assert f0.should_process(fr) == f1.should_process(fr)
        assert np.array_equal(f0.previous_gray, f1.previous_gray) and f0.last_count == f1.last_count

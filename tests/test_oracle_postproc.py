"""Post-processing oracle: PARITY UNPINNED by the reference (dcase_util absent); the median rule is
checked bit-exactly against scipy.ndimage.median_filter (what src/evaluation_measures.py:201 calls)."""
import numpy as np
import pytest
import scipy.ndimage

from oracle import postproc as pp


@pytest.mark.parametrize("win", [1, 2, 3, 7, 14, 15, 27])
@pytest.mark.parametrize("density", [0.1, 0.5, 0.9])
def test_median_rule_equals_scipy(win, density):
    rng = np.random.default_rng(win * 100 + int(density * 10))
    b = (rng.random((313, 20)) < density).astype(np.int64)
    ref = scipy.ndimage.median_filter(b, (win, 1))
    assert np.array_equal(pp.median_filter_time(b, win), ref)


def test_median_short_sequences():
    rng = np.random.default_rng(5)
    for T in (1, 2, 5, 13, 14, 20):
        b = (rng.random((T, 3)) < 0.5).astype(np.int64)
        assert np.array_equal(pp.median_filter_time(b, 14), scipy.ndimage.median_filter(b, (14, 1)))


def test_median_window_constant():
    assert pp.MEDIAN_WINDOW == 14
    assert pp.FRAME_SECONDS == pytest.approx(0.031875)


def test_threshold_is_inclusive():
    p = np.array([[0.5, 0.49999997, 0.50000006]], dtype=np.float32)
    assert pp.binarize(p).tolist() == [[1, 0, 1]]


def test_contiguous_regions_cases():
    assert pp.find_contiguous_regions([0, 0, 0]).tolist() == []
    assert pp.find_contiguous_regions([1, 1, 1]).tolist() == [[0, 3]]
    assert pp.find_contiguous_regions([0, 1, 1, 0, 1]).tolist() == [[1, 3], [4, 5]]
    assert pp.find_contiguous_regions([1, 0, 1, 0]).tolist() == [[0, 1], [2, 3]]
    assert pp.find_contiguous_regions([]).tolist() == []


def test_decode_order_is_class_major():
    m = np.zeros((10, 3), dtype=np.int64)
    m[5:7, 0] = 1
    m[1:3, 2] = 1
    m[0:2, 0] = 1
    assert pp.decode_strong(m) == [(0, 0, 2), (0, 5, 7), (2, 1, 3)]


def test_seconds_and_clip():
    ev = pp.to_seconds([(3, 0, 313), (1, 10, 20)])
    assert ev[0] == (3, 0.0, pytest.approx(313 * 0.031875))
    assert ev[0][2] <= 10.0
    assert ev[1][1] == pytest.approx(0.31875) and ev[1][2] == pytest.approx(0.6375)
    assert pp.to_seconds([(0, 0, 400)])[0][2] == 10.0


def test_encode_strong_frames():
    y = pp.encode_strong([(1.0, 2.0, 4), (9.9, 10.0, 0)])
    assert y.shape == (313, 20)
    on, off = int(1.0 * 32000 // 255 // 4), int(2.0 * 32000 // 255 // 4)
    assert (on, off) == (31, 62)
    assert y[on:off, 4].all() and y[:on, 4].sum() == 0 and y[off:, 4].sum() == 0
    assert y[int(9.9 * 32000 // 255 // 4):313, 0].all()


def test_events_from_strong_end_to_end():
    rng = np.random.default_rng(9)
    strong = rng.random((313, 20)).astype(np.float32)
    strong[100:150, 7] = 0.9
    ev = pp.events_from_strong(strong)
    assert any(c == 7 and on <= 103 and off >= 147 for c, on, off in ev)
    b = scipy.ndimage.median_filter((strong >= 0.5).astype(np.int64), (14, 1))
    assert ev == pp.decode_strong(b)

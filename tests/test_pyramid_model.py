"""Pins the integer model of cv::resize(INTER_LINEAR, 1/2^L) -- the one piece of third-party
arithmetic on the path (find_chessboard_corners.cc:449-450) -- against the real cv2. CPU only."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from oracle import pyoracle as po  # noqa: E402


def _cv(img, level):
    s = 1.0 / (1 << level)
    return cv2.resize(img, None, fx=s, fy=s, interpolation=cv2.INTER_LINEAR)


@pytest.mark.parametrize("level", (1, 2, 3, 4))
def test_oracle_pyramid_equals_cv2_size_sweep(level):
    rng = np.random.default_rng(level)
    for h in list(range(33, 50)) + [480]:
        for w in list(range(40, 73)) + [641, 643]:
            img = rng.integers(0, 256, (h, w), dtype=np.uint8)
            got, want = po.pyramid(img, level), _cv(img, level)
            assert got.shape == want.shape and np.array_equal(got, want), (level, h, w)


def test_oracle_pyramid_equals_cv2_config_sizes():
    rng = np.random.default_rng(7)
    for (w, h) in ((640, 480), (1280, 720), (1920, 1080), (3840, 2160)):
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        for level in (1, 2, 3):
            assert np.array_equal(po.pyramid(img, level), _cv(img, level)), (w, h, level)


@pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref not built")
def test_shim_resize_equals_cv2():
    # the cv::resize stand-in that the reference build (oracle/_ref) is compiled against
    rng = np.random.default_rng(9)
    for level in (1, 2, 3):
        for (w, h) in ((43, 35), (47, 39), (64, 48), (99, 77), (403, 351), (640, 480)):
            img = rng.integers(0, 256, (h, w), dtype=np.uint8)
            assert np.array_equal(po.ref_shim_resize(img, level), _cv(img, level)), (w, h, level)

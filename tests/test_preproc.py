"""The reference CLI's optional --clahe preprocessing (mrgingham-from-image.cc:71-80):
normalize + CLAHE. Oracle pinned against cv2 (CPU); CUDA kernels against the oracle through the C ABI (GPU)."""
import numpy as np
import pytest

from mrgingham_b200 import synth
from oracle import pyoracle as po


def _images():
    rng = np.random.default_rng(9)
    yield synth.board_frame(320, 240, 10, seed=1)
    yield synth.board_frame(403, 351, 10, seed=2)                       # not divisible by the 8x8 tile grid
    yield synth.noise_frame(131, 77, seed=3)
    yield synth.blurred_noise_frame(200, 160, seed=4)
    yield (synth.blob_frame(256, 192, seed=5) // 3 + 40).astype(np.uint8)   # narrow dynamic range
    yield np.full((64, 64), 77, np.uint8)                               # constant: normalises to zero
    for (w, h) in ((8, 8), (9, 17), (16, 9), (33, 65), (640, 480)):
        yield rng.integers(0, 256, (h, w)).astype(np.uint8)


def test_oracle_preproc_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    cl = cv2.createCLAHE(clipLimit=8)
    for img in _images():
        n = cv2.normalize(img, None, 0, 255, cv2.NORM_MINMAX)
        assert np.array_equal(po.normalize_minmax(img), n), img.shape
        assert np.array_equal(po.clahe(img), cl.apply(img)), img.shape
        assert np.array_equal(po.normalize_clahe(img), cl.apply(n)), img.shape

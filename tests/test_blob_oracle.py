"""CPU tests pinning the blob oracle (oracle/blob_oracle.c): against the committed golden vectors
that cv2.SimpleBlobDetector produced (tests/golden/make_blob_golden.py) and, where cv2 is
importable, point-for-point against cv2.findContours on random binaries."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "blobs_v1.npz"))
NAMES = sorted(k[4:] for k in GOLD.files if k.startswith("img/"))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_blobs_match_golden(name):
    got = po.find_blobs(np.ascontiguousarray(GOLD["img/" + name]))
    want = GOLD["pts/" + name]
    assert got.shape == want.shape, (name, got.shape, want.shape)
    assert np.array_equal(got, want), name


def test_oracle_contours_match_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for trial in range(400):
        h, w = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        b = (rng.random((h, w)) < rng.uniform(0.1, 0.9)).astype(np.uint8)
        if trial % 3 == 0:
            b = (cv2.blur(b * 255, (3, 3)) > 100).astype(np.uint8)
        want, _ = cv2.findContours((b * 255).copy(), cv2.RETR_LIST, cv2.CHAIN_APPROX_NONE)
        got = po.blob_find_contours(np.ascontiguousarray(b))
        assert len(want) == len(got), (trial, w, h)
        for cw, cg in zip(want, got):
            assert np.array_equal(cw.reshape(-1, 2), cg), (trial, w, h)


def test_oracle_blobs_on_strided_view():
    img = GOLD["img/circles_small_n7"]
    big = np.zeros((img.shape[0], img.shape[1] + 13), np.uint8)
    big[:, :img.shape[1]] = img
    assert np.array_equal(po.find_blobs(big[:, :img.shape[1]]), GOLD["pts/circles_small_n7"])

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


def test_oracle_blobs_match_cv2_on_random_frames():
    # broader pinning than the 13 committed goldens, where cv2 is importable: the oracle against
    # cv2.SimpleBlobDetector (the reference's parameters) on seeded random frames of several kinds
    cv2 = pytest.importorskip("cv2")
    from mrgingham_b200 import synth
    p = cv2.SimpleBlobDetector_Params()
    p.minArea = 20; p.maxArea = 80000; p.minDistBetweenBlobs = 5; p.blobColor = 0
    det = cv2.SimpleBlobDetector_create(p)
    rng = np.random.default_rng(77)
    nonempty = 0
    for t in range(36):
        w, h = int(rng.integers(40, 420)), int(rng.integers(40, 320))
        kind = t % 4
        if kind == 0:
            img = synth.blob_frame(w, h, seed=1000 + t, nblobs=int(rng.integers(3, 60)))
        elif kind == 1:
            img = synth.circle_grid_frame(max(w, 160), max(h, 160), int(rng.integers(3, 9)), seed=1000 + t,
                                          noise_sigma=float(rng.uniform(0, 6)), blur=bool(t & 4))
        elif kind == 2:
            img = synth.blurred_noise_frame(w, h, seed=1000 + t, passes=int(rng.integers(1, 4)))
        else:
            img = synth.board_frame(max(w, 200), max(h, 160), int(rng.integers(4, 9)), seed=1000 + t)
        kps = det.detect(img)
        pts = np.array([kp.pt for kp in kps], dtype=np.float32).reshape(-1, 2)
        want = np.trunc((pts * np.float32(1000)).astype(np.float64) + 0.5).astype(np.int32)
        got = po.find_blobs(np.ascontiguousarray(img))
        assert got.shape == want.shape and np.array_equal(got, want), (t, kind, img.shape, len(got), len(want))
        nonempty += len(want) > 0
    assert nonempty >= 20

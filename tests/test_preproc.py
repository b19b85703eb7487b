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
    yield synth.blurred_noise_frame(2100, 130, seed=6)                  # several 1024-pixel patches, 17-row tiles
    yield synth.board_frame(1283, 1030, 10, seed=7)
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


@pytest.mark.gpu
def test_gpu_preproc_matches_oracle():
    import torch
    from mrgingham_b200 import api
    api._require_gpu()
    det = api.Detector(max_frames=2)
    for img in _images():
        want = po.normalize_clahe(img)
        got = det.preprocess(img[None], clahe=True, blur_radius=0)[0]
        assert np.array_equal(got, want), img.shape
        got = det.preprocess(img[None], clahe=True, blur_radius=1)[0]
        assert np.array_equal(got, po.box_blur(want, 1)), img.shape
    # a batch (more frames than one chunk), host and device-resident, aligned and odd-offset views
    raw = np.stack([synth.board_frame(403, 351, 10, seed=s) // (1 + s % 3) + 7 * s for s in range(5)]).astype(np.uint8)
    want = np.stack([po.normalize_clahe(f) for f in raw])
    assert np.array_equal(det.preprocess(raw, clahe=True, blur_radius=0), want)
    wide = np.zeros((5, 351, 448), np.uint8)
    for off in (0, 4, 5):
        wide[:, :, off:off + 403] = raw
        t = torch.from_numpy(wide).cuda()[:, :, off:off + 403]
        got = det.preprocess(t, clahe=True, blur_radius=0)
        assert got.is_cuda and np.array_equal(got.cpu().numpy(), want), off
    det.close()


@pytest.mark.gpu
def test_detector_with_clahe_chain():
    # clahe + blur in the detector config = the reference CLI run with --clahe: raw frames in, corners of the
    # equalised, blurred frames out
    from mrgingham_b200 import api
    api._require_gpu()
    raw = np.stack([(synth.board_frame(800, 608, 10, seed=s, blur=False) // 3 + 50).astype(np.uint8) for s in (5, 6, 7)])
    det = api.Detector(max_frames=2, max_points=2048, blur_radius=1, clahe=True)
    for level in (0, 1):
        xy, counts = det.find_corners(raw, level)
        for i in range(len(raw)):
            want = po.find_corners(po.box_blur(po.normalize_clahe(raw[i]), 1), level)
            assert counts[i] == len(want) and len(want) > 0 and np.array_equal(xy[i, :counts[i]], want), (level, i)
    det.close()

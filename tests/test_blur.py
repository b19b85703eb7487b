"""The box blur of the reference CLI's default preprocessing (SURVEY.md row F2,
mrgingham-from-image.cc:106-111): oracle pinned against cv2.blur (CPU), CUDA kernel against the
oracle through the C ABI (GPU)."""
import numpy as np
import pytest

from mrgingham_b200 import synth
from oracle import pyoracle as po


def _images():
    rng = np.random.default_rng(5)
    yield synth.board_frame(320, 240, 10, seed=1, blur=False)
    yield synth.noise_frame(131, 77, seed=2)
    yield (rng.integers(0, 2, (40, 53)) * 255).astype(np.uint8)
    for (w, h) in ((1, 1), (1, 9), (9, 1), (2, 2), (3, 5), (129, 9), (128, 8), (257, 17),
                   (16, 1), (16, 2), (32, 3), (480, 5), (496, 70), (976, 66), (1920, 130)):    # multiples of 16: the fast 3x3 kernel
        yield synth.noise_frame(w, h, seed=10 + w + h)


def test_oracle_blur_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    for img in _images():
        for r in (1, 2, 3):
            want = cv2.blur(img, (1 + 2 * r, 1 + 2 * r))
            assert np.array_equal(po.box_blur(img, r), want), (img.shape, r)


@pytest.mark.gpu
def test_gpu_blur_matches_oracle():
    import torch
    from mrgingham_b200 import api
    api._require_gpu()
    det = api.Detector(max_frames=2)
    for img in _images():
        for r in (1, 2, 4):
            want = po.box_blur(img, r)
            got = det.box_blur(img[None], r)[0]
            assert np.array_equal(got, want), (img.shape, r)
    # device-resident, strided view, two frames, then the detector on the blurred frames (the CLI's default chain)
    raw = np.stack([synth.board_frame(640, 480, 10, seed=s, blur=False) for s in (3, 4)])
    wide = np.zeros((2, 480, 704), np.uint8); wide[:, :, 5:645] = raw
    t = torch.from_numpy(wide).cuda()[:, :, 5:645]
    blurred = det.box_blur(t, 1)
    assert blurred.is_cuda
    xy, counts = det.find_corners(blurred, 0)
    for i in range(2):
        b = po.box_blur(np.ascontiguousarray(raw[i]), 1)
        assert np.array_equal(blurred[i].cpu().numpy(), b)
        want = po.find_corners(b, 0)
        assert counts[i] == len(want) == 100 and np.array_equal(xy[i, :counts[i]], want)
    det.close()


@pytest.mark.gpu
def test_detector_with_cli_default_preprocessing():
    # blur_radius=1 in the detector config = the reference CLI's default chain (blur, then find corners /
    # refine), all on the device: raw frames in, the corners of the blurred frames out
    from mrgingham_b200 import api
    api._require_gpu()
    raw = np.stack([synth.board_frame(800, 608, 10, seed=s, blur=False) for s in (5, 6, 7)])
    det = api.Detector(max_frames=2, max_points=2048, blur_radius=1)
    for level in (0, 1, 2):
        xy, counts = det.find_corners(raw, level)
        for i in range(len(raw)):
            want = po.find_corners(po.box_blur(raw[i], 1), level)
            assert counts[i] == len(want) and np.array_equal(xy[i, :counts[i]], want), (level, i)
    # refinement reads the blurred frame too
    b0 = po.box_blur(raw[0], 1)
    pts = po.find_corners(b0, 1).astype(np.float64) / 1000.0
    lv = np.full(len(pts), 1, np.int8)
    n_want, xy_want, lv_want = po.refine_corners(b0, 0, pts, lv)
    n_got, xy_got, lv_got = det.refine_corners(raw[:1], 0, pts[None], lv[None])
    assert n_got[0] == n_want and np.array_equal(xy_got[0], xy_want) and np.array_equal(lv_got[0], lv_want)
    det.close()
